"""CPU: TF tensor-bundle checkpoints read and written without TensorFlow (drop-in layout)."""
import os

import numpy as np

import chebnet_oracle as O
from multihop_offload_b200 import tf_bundle


def test_crc32c_known_answers():
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283       # CRC-32C check value
    assert tf_bundle.crc32c(b"\x00" * 32) == 0x8A9136AA        # RFC 3720 B.4


def test_reads_shipped_checkpoint(golden_dir):
    d = os.path.join(golden_dir, "ckpt_BAT800")
    prefix = tf_bundle.latest_checkpoint(d)
    assert prefix.endswith("cp-0000.ckpt")
    ws = tf_bundle.load_weights(prefix)
    z = np.load(os.path.join(golden_dir, "weights_BAT800.npz"))
    for i, (W, b) in enumerate(ws):
        np.testing.assert_array_equal(W, z["W%d" % i])
        np.testing.assert_array_equal(b, z["b%d" % i])
    ws_oracle = O.load_reference_weights(d)          # independent reader in the oracle
    for (W, b), (Wo, bo) in zip(ws, ws_oracle):
        np.testing.assert_array_equal(W, Wo); np.testing.assert_array_equal(b, bo)


def test_writer_is_byte_identical_to_tensorflow(golden_dir, tmp_path):
    """Re-saving the shipped weights reproduces TF's own files bit for bit (index, data, state file)."""
    d = os.path.join(golden_dir, "ckpt_BAT800")
    ws = tf_bundle.load_weights(os.path.join(d, "cp-0000.ckpt"))
    out = tmp_path / "model"
    tf_bundle.save_weights(str(out / "cp-0000.ckpt"), ws)
    for name in ("cp-0000.ckpt.index", "cp-0000.ckpt.data-00000-of-00001", "checkpoint"):
        assert (out / name).read_bytes() == open(os.path.join(d, name), "rb").read(), name


def test_roundtrip_other_orders(tmp_path):
    rng = np.random.default_rng(0)
    for K in (1, 3, 5):
        ws = O.glorot_weights([4, 32, 32, 32, 32, 1], K, rng)
        ws = [(W, rng.normal(size=b.shape)) for W, b in ws]
        p = tf_bundle.save_weights(str(tmp_path / ("k%d" % K) / "cp-0007.ckpt"), ws)
        assert tf_bundle.latest_checkpoint(os.path.dirname(p)) == p
        back = tf_bundle.load_weights(p)
        for (W, b), (W2, b2) in zip(ws, back):
            np.testing.assert_array_equal(W, W2); np.testing.assert_array_equal(b, b2)


def test_missing_checkpoint_dir(tmp_path):
    assert tf_bundle.latest_checkpoint(str(tmp_path)) is None
