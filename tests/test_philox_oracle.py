"""Known-answer tests for the Philox4x32-10 restatement (Random123 kat_vectors)."""
import numpy as np

from oracle import philox


def _run(c, k):
    return philox.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32)).tolist()


def test_random123_known_answers():
    assert _run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _run([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_open_interval_and_moments():
    x = np.array([0, 0xFF, 0xFFFFFFFF], dtype=np.uint32)
    u = philox.u32_to_uniform(x)
    assert u[0] == u[1] == np.float32(0.5 * 2.0 ** -23) and 0 < u[2] < 1
    z = philox.standard_normal(seed=7, iteration=1, horizon=30, n=2048, act_dim=6)
    assert z.shape == (30 * 2048, 6)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


def test_shard_independence():
    """Global candidate index is the counter word, so a shard starting at offset 96 draws
    exactly rows 96.. of the unsharded population."""
    full = philox.standard_normal(3, 2, 5, 128, 6).reshape(5, 128, 6)
    part = philox.standard_normal(3, 2, 5, 32, 6, cand_offset=96).reshape(5, 32, 6)
    np.testing.assert_array_equal(full[:, 96:], part)
