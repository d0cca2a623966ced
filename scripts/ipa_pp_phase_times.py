"""Per-phase clock64 timeline of the ping-pong pass 1 (ipa_tc_pp.cu), one record per item (developer diagnostics)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from debug_ipa_tc_common import make, split, ops, dev, H
from se3diff_b200 import _lib
B, Lm = int(os.environ.get('IPA_B', 256)), int(os.environ.get('IPA_L', 84))
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
sc_, pt_ = split(proj)
both_ = torch.cat([sc_, pt_.to(torch.bfloat16)], dim=1)
sc_, pt_ = both_[:, :sc_.shape[1]], both_[:, sc_.shape[1]:]
run = lambda: ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out=out)
for _ in range(3): run()
n_items = B * H
buf = torch.zeros(n_items * 32, dtype=torch.int64, device=dev)
lib = _lib.lib(); lib.se3_debug_set_pp_phase_buffer.argtypes = [C.c_void_p]; lib.se3_debug_set_pp_phase_buffer.restype = None
lib.se3_debug_set_pp_phase_buffer(C.c_void_p(buf.data_ptr())); run(); torch.cuda.synchronize(); lib.se3_debug_set_pp_phase_buffer(None)
full = buf.view(-1, 32).double().cpu()
t = full[:, :8]
d = t[:, 1:] - t[:, :-1]
names = ["wait inputs", "transform", "wait turn", "pass A (+bias wait, row max)", "pass B", "wait P.V", "epilogue"]
print("mean cycles per phase:", {n: round(v) for n, v in zip(names, d.mean(0).tolist())}, "item total", round((t[:, 7] - t[:, 0]).mean().item()))
print("p10:", {n: round(v) for n, v in zip(names, d.quantile(0.1, dim=0).tolist())})
print("p90:", {n: round(v) for n, v in zip(names, d.quantile(0.9, dim=0).tolist())})
sms = torch.cuda.get_device_properties(0).multi_processor_count
stride = 2 * min(sms, (n_items + 1) // 2)
w0 = t[0::stride]          # the items of worker 0, in order
print("worker 0, first items: start-to-start", [round(v) for v in (w0[1:6, 0] - w0[:5, 0]).tolist()])
print("worker 0 item 2 stamps rel.:", [round(v) for v in (w0[2] - w0[2, 0]).tolist()], " worker 1 item 2:", [round(v) for v in (t[1::stride][2] - w0[2, 0]).tolist()])

rel = full - full[:, :1]
k = 2 * stride + 0
print("item", k, "consumer chunk-ready stamps (rel. to item start):", [round(v) for v in rel[k, 8:16].tolist()])
print("item", k, "issuer: ops_ready seen", round(rel[k, 24].item()), " chunk issue times", [round(v) for v in rel[k, 16:24].tolist()], " p_ready seen", round(rel[k, 25].item()), " P.V committed", round(rel[k, 26].item()))
print("mean over items: ops_ready seen", round(rel[:, 24].mean().item()), "chunk-ready", [round(v) for v in rel[:, 8:14].mean(0).tolist()], "issue", [round(v) for v in rel[:, 18:22].mean(0).tolist()],
      "p_ready seen", round(rel[:, 25].mean().item()), "P.V committed", round(rel[:, 26].mean().item()), "consumer stamps", [round(v) for v in rel[:, :8].mean(0).tolist()])
print("mean: group 0 chunk 0: ready", round(rel[:, 8].mean().item()), "loaded", round(rel[:, 27].mean().item()), "math+store done", round(rel[:, 28].mean().item()),
      "| issuer chunk 2: issue start", round(rel[:, 18].mean().item()), "committed", round(rel[:, 30].mean().item()), "| group 0 sees chunk 2", round(rel[:, 10].mean().item()))
