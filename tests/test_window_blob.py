"""Wire / on-disk format of a window (include/vilba.h: vilba_window_serialize / _deserialize): host code, no GPU.

The golden blob pins the format byte for byte: a change of the layout must bump the version and regenerate
tests/golden/window_tiny_v1.blob (tests/golden/make_window_blob.py)."""
import hashlib
import os

import numpy as np
import pytest

from mc_slam_b200 import capi, synth
from mc_slam_b200.capi import Window

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "window_tiny_v1.blob")
FIELDS = ["kf_state", "kf_flags", "kf_id", "imu_kf_i", "imu_kf_j", "imu_preint", "pt_xyz", "pt_obs_begin", "obs_kf", "obs_uv",
          "obs_inv_sigma2", "Rbc", "Pbc", "gravity"]


def _same(a: Window, b: Window):
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), f
    assert (a.fx, a.fy, a.cx, a.cy) == (b.fx, b.fy, b.cx, b.cy)


@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_round_trip_is_exact(name):
    w = synth.make_config(name, n_fixed_extra=1)
    blob = w.to_bytes()
    assert len(blob) % 8 == 0 and blob[:8] == b"VILBAWIN"
    _same(w, Window.from_bytes(blob))
    assert Window.from_bytes(blob).to_bytes() == blob  # canonical: padding bytes are zeros


def test_window_without_points_round_trips():
    import dataclasses
    w = synth.make_config("tiny")
    w0 = dataclasses.replace(w, pt_xyz=np.zeros((0, 3)), pt_obs_begin=np.zeros(1, np.int32), obs_kf=np.zeros(0, np.int32),
                             obs_uv=np.zeros((0, 2), np.float32), obs_inv_sigma2=np.zeros(0, np.float32), truth={})
    _same(w0, Window.from_bytes(w0.to_bytes()))


def test_golden_blob_pins_the_format():
    w = synth.make_config("tiny")
    blob = w.to_bytes()
    with open(GOLDEN, "rb") as f:
        gold = f.read()
    assert hashlib.sha256(blob).hexdigest() == hashlib.sha256(gold).hexdigest()
    _same(w, Window.from_bytes(gold))


def test_corruption_truncation_and_foreign_data_are_rejected():
    blob = bytearray(synth.make_config("tiny").to_bytes())
    ok = bytes(blob)
    Window.from_bytes(ok)
    for pos in (300, len(blob) // 2, len(blob) - 1):  # payload bit flips: checksum
        bad = bytearray(ok)
        bad[pos] ^= 0x10
        with pytest.raises(ValueError):
            Window.from_bytes(bytes(bad))
    for bad in (ok[:100], ok[:-8], b"NOTAWINDOW" + ok[10:], ok[:8] + (2).to_bytes(4, "little") + ok[12:]):
        with pytest.raises(ValueError):
            Window.from_bytes(bad)
    # a count that does not match the payload length
    bad = bytearray(ok)
    bad[16:20] = (10 ** 6).to_bytes(4, "little")
    with pytest.raises(ValueError):
        Window.from_bytes(bytes(bad))


def test_out_of_range_indices_are_rejected_even_with_a_valid_checksum():
    w = synth.make_config("tiny")
    w.obs_kf = w.obs_kf.copy()
    w.obs_kf[3] = 77  # the serialiser does not validate; the reader does
    with pytest.raises(ValueError):
        Window.from_bytes(w.to_bytes())


def test_blob_size_matches_the_layout():
    w = synth.make_config("small")
    pad8 = lambda n: (n + 7) // 8 * 8
    K, NI, P, E = w.n_kf, w.n_imu, w.n_pts, w.n_obs
    expect = 256 + sum(pad8(x) for x in (8 * 22 * K, K, 8 * K, 4 * NI, 4 * NI, 8 * 142 * NI, 24 * P, 4 * (P + 1), 4 * E, 8 * E, 4 * E))
    assert len(w.to_bytes()) == expect


@pytest.mark.gpu
def test_replayed_blob_solves_like_the_original(vilba, oracle):
    from parity_util import compare
    w = synth.make_config("small", window_index=4)
    w2 = Window.from_bytes(w.to_bytes())
    c = vilba.Context(0)
    try:
        r1, r2 = c.local_ba(w), c.local_ba(w2)
        assert np.array_equal(r1.kf_state, r2.kf_state) and np.array_equal(r1.pt_xyz, r2.pt_xyz)
        compare(r2, oracle.local_ba(w), w)
    finally:
        c.close()
