"""Timing of the batched captcha client-id kernel on the config-2 request stream."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import torch  # noqa: E402

import synth  # noqa: E402
from pingoo_b200 import Action, Rule, WafEngine  # noqa: E402

batch = synth.RequestStream(config_id=2, payloads=[]).generate(0, 1_000_000)
eng = WafEngine([Rule("r", None, [Action.BLOCK])], device=0)
t, cb = eng.to_device(batch)
out = torch.zeros((batch.n, 44), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    eng.client_ids_device(cb, out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.client_ids_device(cb, out, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
by = batch.total["user_agent"] + batch.total["host"] + batch.n * (17 + 8 + 44)
print(f"client ids: {ms:.3f} ms per 1M requests  {batch.n / ms / 1e3:.0f} M ids/s  {by / ms / 1e6:.0f} GB/s algorithmic")
