import os, sys, torch
sys.path.insert(0, "/root/repo")
from uc2_b200._lib import call, stream, lib
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for B, S in [(400, 30), (400, 48), (400, 64), (400, 80), (400, 100), (400, 118), (400, 129), (64, 160), (48, 222)]:
    qkv = torch.randn(B * S, 2304, device="cuda").bfloat16()
    lens = torch.randint(max(1, S * 2 // 3), S + 1, (B,), device="cuda"); lens[0] = S
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).long().contiguous()
    ctx = torch.empty(B * S, 768, dtype=torch.bfloat16, device="cuda"); lse = torch.empty(B, 12, S, device="cuda")
    f = lambda: call("uc2_attention_fwd", qkv.data_ptr(), mask.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, S, stream())
    lib().uc2_attention_tc_enable(1); a = t(f)
    lib().uc2_attention_tc_enable(0); b = t(f)
    lib().uc2_attention_tc_enable(1)
    print(f"B={B} S={S}: tcgen05 {a:.1f} us, mma.sync {b:.1f} us")
