"""Per-phase clock64 timeline of the pass-1 CTAs (developer diagnostics)."""
import sys, os, math, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from debug_ipa_tc_common import make, head_major, split, ops, dev, H
from se3diff_b200 import _lib
B, Lm = int(os.environ.get('IPA_B', 256)), int(os.environ.get('IPA_L', 84))
n_cta = B * H * ((Lm + 127) // 128) * (2 if Lm > 256 else 1)
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
sc_, pt_ = split(proj)
both_ = torch.cat([sc_, pt_.to(torch.bfloat16)], dim=1)      # production layout: one projection writes scalar | bf16 point records
sc_, pt_ = both_[:, :sc_.shape[1]], both_[:, sc_.shape[1]:]
run = lambda: ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out=out)
for _ in range(3): run()
buf = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
lib = _lib.lib(); lib.se3_debug_set_phase_buffer.argtypes = [C.c_void_p]; lib.se3_debug_set_phase_buffer.restype = None
lib.se3_debug_set_phase_buffer(C.c_void_p(buf.data_ptr())); run(); torch.cuda.synchronize(); lib.se3_debug_set_phase_buffer(None)
t = buf.view(-1, 16).double().cpu()
st = t[:, [0, 8, 9, 10, 11, 12, 1]]
ds = st[:, 1:] - st[:, :-1]
print('staging detail:', {n: round(v) for n, v in zip(['alloc', 'TMA issue + key bias', 'cp.async wait', 'CTA barrier + TMA landing', 'frame transforms', 'fences'], ds.mean(0).tolist())})
t = t[:, :8]
d = t[:, 1:] - t[:, :-1]
names = ["staging(+alloc)", "sync+MMA1+wait", "passA(+bias wait)", "passB", "sync+MMA2+wait", "epilogue", "dealloc"]
print("mean cycles per phase:", {n: round(v) for n, v in zip(names, d.mean(0).tolist())}, "total", round((t[:, 7] - t[:, 0]).mean().item()))
print("p90:", {n: round(v) for n, v in zip(names, d.quantile(0.9, dim=0).tolist())})
# persistent editions: cycles from a CTA's first stamp to its last one = the kernel's duration in SM cycles (against the event
# time of the same launch this gives the clock the SMs actually ran at)
G = int(os.environ.get('IPA_GRID', 0))
if G:
    tt = buf.view(-1, 16).double().cpu()
    first = tt[:G, 0]
    last = torch.stack([tt[c::G, 7].max() for c in range(G)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    span = (last - first)
    print(f"persistent CTAs: span cycles mean {span.mean():.0f} max {span.max():.0f}; pass 1 + pass 2 event time {ms * 1e3:.1f} us")
    gt = torch.stack([tt[c::G, 14].max() - tt[c::G, 14].min() for c in range(G)])       # ns between the CTA's first and last item end
    cy = torch.stack([tt[c::G, 7].max() - tt[c::G, 7].min() for c in range(G)])
    print(f"SM clock inside the kernel: {(cy / gt).mean():.3f} GHz (min {(cy / gt).min():.3f}, max {(cy / gt).max():.3f}); kernel wall span {(tt[:, 14].max() - tt[:, 14].min()) / 1e3:.1f} us from the first item end to the last")
