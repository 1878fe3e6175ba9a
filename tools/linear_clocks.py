"""SM-clock timeline of one k_lin CTA (tuning)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flatland_marl_b200.policy import BatchedActor
actor = BatchedActor(None, seed=0)
dev = actor.device
for (M, K, N, act) in [(51200, 256, 256, 1)]:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    clk = torch.zeros(128, dtype=torch.int64, device=dev)
    for _ in range(3):
        actor.lib.fl_policy_linear_debug(a.data_ptr(), K, w.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, act, clk.data_ptr(), None)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    actor.lib.fl_policy_linear_debug(a.data_ptr(), K, w.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, act, clk.data_ptr(), None)
    e.record()
    torch.cuda.synchronize()
    c = clk.cpu().numpy()
    t0 = c[3]
    print("M=%d K=%d N=%d act=%d: %.1f us; kernel start->mma warp %d, weights resident %d, end %d cycles" % (M, K, N, act, s.elapsed_time(e) * 1e3, c[0] - t0, c[1] - t0, c[2] - t0))
    for t in range(8):
        if c[8 + 4 * t] == 0:
            break
        print("  tile %d: producer start %6d | mma: acc free %6d first full %6d last full %6d | epi: acc full %6d read %6d stored %6d" %
              (t, c[80 + t] - t0, c[8 + 4 * t] - t0, c[9 + 4 * t] - t0, c[10 + 4 * t] - t0, c[48 + 4 * t] - t0, c[49 + 4 * t] - t0, c[50 + 4 * t] - t0))
        print("          epilogue detail: math done %6d, staged %6d, bar2 %6d, (unused) %6d" % tuple(c[96 + 4 * t + i] - t0 for i in range(4)))
