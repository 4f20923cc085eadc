"""Development aid: isolates HostPipeline / graph-replay determinism (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.init import spread_state_dict
from spegnet_b200 import SPEGNet, HostPipeline

CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}
sd = spread_state_dict(0)
model = SPEGNet(CFG); model.load_state_dict(sd); model = model.cuda().eval()
g = torch.Generator().manual_seed(33)

def run(x):
    with torch.no_grad():
        o = model(x.cuda())
    return o["predictions"][-1].cpu(), o["edge"].cpu()

for B in (1, 2, 3, 4, 5):
    x = torch.randn(B, 3, 256, 256, generator=g)
    a = run(x); b = run(x); c = run(x)
    model.cuda_graph_max_batch = 0
    e = run(x); f = run(x)
    model.cuda_graph_max_batch = 8
    print(f"B={B} graph==graph {torch.equal(a[0], b[0]) and torch.equal(b[0], c[0])} eager==eager {torch.equal(e[0], f[0])} "
          f"graph==eager {torch.equal(a[0], e[0])} maxdiff {(a[0]-e[0]).abs().max().item():.3e}", flush=True)

def pipe_case(name, sizes, depth, warm):
    batches = [torch.randn(b, 3, 256, 256, generator=g).pin_memory() for b in sizes]
    want = [run(x) for x in batches]
    if warm:
        with torch.no_grad():
            model(torch.randn(16, 3, 256, 256, generator=g).cuda())
    got = [(o["prediction"].clone(), o["edge"].clone()) for o in HostPipeline(model, depth=depth).run(batches)]
    ok = [torch.equal(gp, wp) and torch.equal(ge, we) for (gp, ge), (wp, we) in zip(got, want)]
    print(name, sizes, "depth", depth, "warm", warm, ok, flush=True)

pipe_case("uniform", (2, 2, 2, 2, 2), 2, False)
pipe_case("uniform", (2, 2, 2, 2, 2), 3, False)
pipe_case("uniform3", (3, 3, 3, 3, 3), 3, False)
pipe_case("ragged", (3, 3, 2, 3, 1, 2), 2, False)
pipe_case("ragged", (3, 3, 2, 3, 1, 2), 3, False)
pipe_case("ragged", (3, 3, 2, 3, 1, 2), 3, True)
pipe_case("ragged", (2, 2, 2, 2, 1), 3, True)
