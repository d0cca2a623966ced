"""Times the tensor-core IPA operator (pass 1 + pass 2) at the bench shape with CUDA events (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from debug_ipa_tc_common import make, split, ops, dev, H
B, Lm = int(os.environ.get("IPA_B", 256)), int(os.environ.get("IPA_L", 84))
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
sc_, pt_ = split(proj)
both_ = torch.cat([sc_, pt_.to(torch.bfloat16)], dim=1)      # production layout: one projection writes scalar | bf16 point records
sc_, pt_ = both_[:, :sc_.shape[1]], both_[:, sc_.shape[1]:]
run = lambda: ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out=out)
for _ in range(5): run()
torch.cuda.synchronize()
n = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('SE3DIFF_B200_LIB', 'default')}: B={B} L={Lm}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call (pass 1 + pass 2), checksum {float(out.float().abs().mean()):.6f}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): run()
    torch.cuda.synchronize()
for r in sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:3]:
    print(f"   {r.device_time_total / r.count:9.1f} us  {r.key[:90]}")
