import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from test_gpu_trainers import _cnn32_like
from eoe_b200.training import TRAINER
from eoe_b200.training import ad_trainer as AT
dev = "cuda"
g = torch.Generator().manual_seed(0)
loader = [(torch.randn(256, 3, 32, 32, generator=g).to(dev), (torch.arange(256) >= 128).long().to(dev), None) for _ in range(40)]
orig = AT.ADTrainer._capture_step
ncap = [0]
def counted(self, *a, **k):
    ncap[0] += 1
    t0 = time.perf_counter(); r = orig(self, *a, **k); torch.cuda.synchronize(); print("capture took ms", (time.perf_counter() - t0) * 1e3, "lr obj", type(r[5]))
    return r
AT.ADTrainer._capture_step = counted
for sgd in (False, True):
    for graph in (False, True):
        model = _cnn32_like()
        tr = TRAINER["hsc"](model, epochs=1, lr=1e-3, device=dev, graph_step=graph, sgd=sgd)
        tr.train_cls(model, loader[:8], nominal_label=0)
        torch.cuda.synchronize()
        for ep in (1, 2, 4):
            ncap[0] = 0
            t0 = time.perf_counter()
            tr.train_cls(model, loader, nominal_label=0, epochs=ep)
            torch.cuda.synchronize()
            print(dict(sgd=sgd, graph=graph, epochs=ep, ms_total=(time.perf_counter() - t0) * 1e3, captures=ncap[0]))
