"""Ensemble streaming (sllb_sim4d_stream_step): upload of the next state and download of the previous one overlap the
current state's time step; the results must be bit-identical to upload -> run(1) -> download, state by state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import selalib_b200 as s
    s.init(0)
    return s


def test_stream_step_equals_serial_steps(sb):
    import torch
    nc = [32, 32, 32, 32]
    args = (nc, [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-2, 0.1)
    S = sb.Sim4d(*args)
    f0 = S.field().download()
    rng = np.random.default_rng(20261017)
    members = [np.asfortranarray(f0 * (1.0 + 0.05 * k) + 1e-3 * rng.standard_normal(f0.shape)) for k in range(4)]
    serial = []
    for m in members:
        S.field().upload(m)
        S.run(1, diagnostics=False)
        serial.append(S.field().download())
    n = f0.size
    hin = [torch.from_numpy(np.ascontiguousarray(m.reshape(-1, order="F"))).pin_memory() for m in members]
    hout = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in members]
    # N members take N + 2 calls
    for k in range(len(members) + 2):
        nxt = hin[k].data_ptr() if k < len(members) else None
        prv = hout[k - 2].data_ptr() if k >= 2 else None
        S.stream_step(nxt, prv)
    for k, ref in enumerate(serial):
        got = hout[k].numpy().reshape(f0.shape, order="F")
        assert np.array_equal(got, ref), k
    # the object is still usable as an ordinary simulation afterwards
    S.field().upload(members[0])
    S.run(1, diagnostics=False)
    assert np.array_equal(S.field().download(), serial[0])
    S.destroy()
