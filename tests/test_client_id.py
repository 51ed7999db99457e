"""Captcha client id for a batch: generate_captcha_client_id (pingoo/captcha.rs:409-421) =
base64url-no-pad(SHA-256(ip octets || user_agent || host)).  The checker is hashlib + base64 (standard algorithms; the
composition is the reference's), so this row's parity is pinned."""
import base64
import hashlib

import numpy as np
import pytest

from pingoo_b200 import WafEngine, Rule, Action, pack_requests


def client_id_ref(ip16: bytes, is_v6: int, ua: bytes, host: bytes) -> bytes:
    h = hashlib.sha256()
    h.update(ip16 if is_v6 else ip16[:4])
    h.update(ua)
    h.update(host)
    return base64.urlsafe_b64encode(h.digest()).rstrip(b"=")


def test_reference_composition_known_answer():
    # sha256(7f000001 || "Mozilla/5.0" || "example.com")
    want = base64.urlsafe_b64encode(hashlib.sha256(bytes([127, 0, 0, 1]) + b"Mozilla/5.0" + b"example.com").digest()).rstrip(b"=")
    assert client_id_ref(bytes([127, 0, 0, 1]) + bytes(12), 0, b"Mozilla/5.0", b"example.com") == want and len(want) == 43


@pytest.mark.gpu
def test_client_ids_match_hashlib():
    import torch

    reqs = []
    # message lengths around the SHA-256 block boundaries (55/56/64/119/120 bytes), empty strings, IPv6, maximum sizes
    for total in [0, 1, 50, 51, 52, 55, 56, 59, 60, 63, 64, 115, 116, 119, 120, 127, 128, 300, 512]:
        for v6 in (False, True):
            ipl = 16 if v6 else 4
            body = max(0, total - ipl)
            ual = min(256, body // 2 + body % 2)
            hol = min(256, body - ual)
            reqs.append(dict(host="h" * hol, url="/", path="/", method="GET", user_agent=("u" * ual) if ual else "",
                             ip="2001:db8::%x" % (total + 1) if v6 else "10.0.%d.%d" % (total // 256, total % 256), remote_port=1))
    rng = np.random.RandomState(3)
    for i in range(2000):
        ua = "".join(chr(32 + int(c)) for c in rng.randint(0, 95, size=rng.randint(0, 257))).strip()
        host = "".join(chr(97 + int(c)) for c in rng.randint(0, 26, size=rng.randint(0, 257)))
        reqs.append(dict(host=host, url="/", path="/", method="GET", user_agent=ua, ip="192.168.%d.%d" % (i // 256, i % 256) if i % 3 else "fe80::%x" % i, remote_port=1))
    batch = pack_requests(reqs)
    eng = WafEngine([Rule("r", None, [Action.BLOCK])], device=0)
    t, cb = eng.to_device(batch)
    out = torch.zeros((batch.n, 44), dtype=torch.uint8, device="cuda")
    eng.client_ids_device(cb, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for i in range(batch.n):
        want = client_id_ref(bytes(batch.ip[i]), int(batch.ip_is_v6[i]), batch.field("user_agent", i), batch.field("host", i))
        assert bytes(got[i, :43]) == want and got[i, 43] == 0, f"request {i}: ua {len(batch.field('user_agent', i))} B host {len(batch.field('host', i))} B"
