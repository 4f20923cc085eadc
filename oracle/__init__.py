"""CPU oracle for the SPEGNet inference forward pass -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import anything from this package.  The product path (``spegnet_b200``) never does and
fails loudly when its CUDA library is missing.

Contents
--------
``hiera``        plain-PyTorch fp32 restatement of the SAM2 Hiera trunk (third-party ``sam2`` package,
                 unpinned git HEAD, absent from /root/reference -- **parity unpinned** upstream; pinned
                 here against the independent HF port ``transformers.models.sam2``).
``head``         functional fp32 restatement of CFI / EFE / PED (models/feature_integration.py,
                 models/object_detection.py, models/spegnet.py) -- pinned against the reference modules
                 imported verbatim (tests/golden/make_golden.py, run in the dev container).
``spegnet``      the whole forward = trunk + head, state-dict driven, reference key names.
``sam2_shim``    ``sys.modules`` shim that lets the reference ``models/spegnet.py`` construct verbatim.
``init``         seeded "spread" random init in the reference checkpoint schema.
``sod_metrics``  numpy/scipy restatement of the ``py_sod_metrics`` scores used by utils/metrics.py
                 (third-party ``pysodmetrics``, unpinned, absent -- **parity unpinned**; pinned only by
                 analytic known-answer tests).
``preprocess``   CODImageProcessor.process_image from the decoded RGB array on (utils/image_processor.py:114-134) and
                 the prediction resize of engine/predictor.py:350-365 -- pinned bit-exact against the reference
                 class run on PNG fixtures (tests/golden/make_golden_preprocess.py -> preprocess.npz).
"""
