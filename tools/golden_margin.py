"""Max |delta sigmoid| of the four outputs against the reference golden file and the oracle (development aid; GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.init import spread_state_dict
from spegnet_b200 import SPEGNet
CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}
sd = spread_state_dict(0)
m = SPEGNet(CFG); m.load_state_dict(sd); m = m.cuda().eval()
gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "full_512.npz"))
def err(a, b): return float((a.float().cpu().sigmoid() - torch.as_tensor(np.asarray(b, dtype=np.float32)).sigmoid()).abs().max())
for seed in [int(gold["input_seed"]), 11, 12, 13]:
    x = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad(): out = m(x.cuda())
    if seed == int(gold["input_seed"]):
        print("golden seed", seed, [round(err(out["predictions"][i], gold[f"pred{i+1}"]), 5) for i in range(3)], round(err(out["edge"], gold["edge"]), 5))
    else:
        from oracle.spegnet import spegnet_forward
        ref = spegnet_forward(sd, x)
        print("oracle seed", seed, [round(err(out["predictions"][i], ref["predictions"][i]), 5) for i in range(3)], round(err(out["edge"], ref["edge"]), 5))
