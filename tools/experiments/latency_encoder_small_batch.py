import sys, time, torch
sys.path.insert(0, '/root/repo')
from eoe_b200.encoder import ClipImageEncoder
from eoe_b200.synth import random_vit_state_dict
dev = torch.device('cuda', 0)
for P, Bs in ((32, (16, 64, 128, 256)), (16, (16, 64, 128))):
    sd = random_vit_state_dict(P, seed=0)
    for B in Bs:
        enc = ClipImageEncoder(sd, device=dev, max_batch=B)
        imgs = torch.randn(B, 3, 224, 224, device=dev)
        text = torch.nn.functional.normalize(torch.randn(10, 512, device=dev), dim=-1)
        out = torch.empty(B, device=dev)
        for _ in range(5): enc.score(imgs, text, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(50): enc.score(imgs, text, out=out)
        e1.record(); t_cpu = time.perf_counter() - t0
        torch.cuda.synchronize(); t_wall = time.perf_counter() - t0
        # graph replay
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            enc.score(imgs, text, out=out)
            torch.cuda.synchronize()
            try:
                with torch.cuda.graph(g, stream=s):
                    enc.score(imgs, text, out=out)
                ok = True
            except Exception as ex:
                ok = False; print('graph capture failed', ex)
        tg = None
        if ok:
            for _ in range(3): g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50): g.replay()
            b.record(); torch.cuda.synchronize()
            tg = a.elapsed_time(b) / 50
        print(f"P={P} B={B}: eager device {e0.elapsed_time(e1)/50:.3f} ms/call, cpu enqueue {t_cpu/50*1e3:.3f} ms/call, wall {t_wall/50*1e3:.3f}; graph replay {tg} ms/call", flush=True)
        del enc
