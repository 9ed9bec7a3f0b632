import sys, os, torch
sys.path.insert(0, "simple-diffusion-model_b200"); sys.path.insert(0, "tools")
from b200 import ops
from bench_gemm_layers import timeit
for n, c, hw in ((256, 1024, 16), (256, 128, 64), (256, 512, 32), (256, 256, 32), (256, 512, 8)):
    p = hw * hw
    sets = []
    for _ in range(3):
        y = torch.randn((n, hw, hw, c), device="cuda", dtype=torch.bfloat16)
        out = torch.empty_like(y)
        st = torch.stack([torch.zeros((n, 32), device="cuda"), torch.full((n, 32), float(p * c // 32), device="cuda")], dim=-1).contiguous()
        sets.append((y, out, st, torch.ones(c, device="cuda"), torch.zeros(c, device="cuda"), torch.randn((1, c), device="cuda")))
    t = timeit(lambda y, out, st, ga, be, s: ops.adagn_apply(y, st, ga, be, s, 0, out=out), sets)
    t2 = timeit(lambda y, out, st, ga, be, s: ops.adagn_apply(y, st, ga, be, s, 0, out=out, residual=y), sets)
    e = n * p * c
    print(f"N{n} C{c} {hw}x{hw}: plain {t*1e3:7.1f} us {e*4/t/1e6:7.0f} GB/s | +res {t2*1e3:7.1f} us {e*6/t2/1e6:7.0f} GB/s")
