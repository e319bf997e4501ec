"""Times ClipAdam.step() alone on the C1 / C3 parameter set (gradients filled with noise)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, bench
    from videocad_b200 import AutoRegressiveTransformer
    from videocad_b200.optim import ClipAdam
    cfg = bench.CONFIGS[sys.argv[2]]
    dev = torch.device("cuda", 0)
    model = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", **cfg["model"]).to(dev)
    ps = list(model.parameters())
    opt = ClipAdam(ps, lr=1e-5, max_norm=1.0)
    n = sum(p.numel() for p in ps)
    for p in ps:
        p.grad = torch.randn_like(p) * 1e-3
    for _ in range(5):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 50
    e0.record()
    for _ in range(K):
        opt.step()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / K
    print(f"{sys.argv[2]} params={n/1e6:.1f}M  {us:.1f} us/step  "
          f"{9*4*n/us/1e3:.0f} GB/s (9 x 4 B per parameter)  norm={float(opt.last_grad_norm):.6f}")
else:
    for cfgname in ("c1", "c3"):
        subprocess.run([sys.executable, __file__, "child", cfgname])
