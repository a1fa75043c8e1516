import torch, time, sys
sys.path.insert(0, ".")
import monotonic_rnnt_b200 as mr
wl = mr.synth.workload("c3")
B, Tm, U, V = wl.B, int(wl.T.max()), int(wl.S.max()) + 1, wl.V
x = torch.rand(B, Tm, U, V, device="cuda")
T = torch.from_numpy(wl.T).cuda(); S = torch.from_numpy(wl.S).cuda()
t = torch.arange(Tm, device="cuda")[None, :, None]; u = torch.arange(U, device="cuda")[None, None, :]
mask = (t < T[:, None, None]) & (u <= S[:, None, None])
idx = mask.reshape(-1).nonzero().squeeze(1)
flat = x.reshape(-1, V)
for _ in range(3):
    packed = flat.index_select(0, idx); g = torch.zeros_like(flat); g.index_copy_(0, idx, packed)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
e0.record(); packed = flat.index_select(0, idx); e1.record(); g = torch.zeros_like(flat); g.index_copy_(0, idx, packed); e2.record(); torch.cuda.synchronize()
print("caller-side gather %.1f us, zero+scatter of gradients %.1f us (torch index_select / index_copy_)" % (e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
