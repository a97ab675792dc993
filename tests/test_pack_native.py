"""The library's native packer (pf_pack_plan / pf_pack_2bit / pf_pack_4bit) against the numpy
reference packing in panfeed_b200/packer.py.  Pure host code: runs without a GPU."""
import numpy as np
import pytest

from panfeed_b200 import capi, packer


def _seqs(rng, n, amb_rate, max_len=400):
    out = []
    for _ in range(n):
        L = int(rng.integers(0, max_len))
        s = rng.choice(list("ACGT"), L)
        if amb_rate and rng.random() < 0.3:
            m = rng.random(L) < amb_rate
            s[m] = rng.choice(list("NRYKMSWBDHVX"), int(m.sum()))
        out.append("".join(s).encode())
    return out


@pytest.mark.parametrize("amb_rate,threads", [(0.0, 1), (0.0, 4), (0.02, 1), (0.02, 4)])
def test_native_packer_matches_numpy(amb_rate, threads):
    rng = np.random.default_rng(int(amb_rate * 1000) + threads)
    seqs = _seqs(rng, 9000 if threads > 1 else 500, amb_rate)
    got = capi.pack_sequences(seqs, n_threads=threads)
    want = packer.pack_planes_numpy(seqs)
    assert np.array_equal(got[0], want[0])            # 2-bit plane
    assert np.array_equal(got[1], want[1])            # base offsets
    assert np.array_equal(got[2], want[2])            # ambiguous flags
    if want[3] is None:
        assert got[3] is None
    else:
        assert np.array_equal(got[3], want[3])        # 4-bit plane
    assert np.array_equal(got[4], want[4])            # offsets in the 4-bit plane


def test_native_packer_edge_cases():
    got = capi.pack_sequences([])
    assert len(got[0]) == 0 and got[3] is None
    got = capi.pack_sequences([b"", b"A" * 64, b"ACGT" * 16 + b"T"])
    want = packer.pack_planes_numpy([b"", b"A" * 64, b"ACGT" * 16 + b"T"])
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    with pytest.raises(ValueError):
        capi.pack_sequences([b"ACGTZ"])
    with pytest.raises(ValueError):
        packer.pack_planes_numpy([b"ACGTZ"])
