"""CPU: the reference's WarpFrame (src/utils.py:27-31) is cv2.resize(frame, (84, 84), INTER_AREA).
For the boards whose padded edge divides 84 (V = 12 -> x7, V = 21 -> x4) that is EXACT pixel
replication, which is what obs_mode='atari84' implements on the device."""
import numpy as np
import pytest

import snake_oracle as so


@pytest.mark.parametrize("D,S", [(10, 2), (10, 3), (19, 2), (19, 3)])
def test_inter_area_is_exact_replication(D, S):
    cv2 = pytest.importorskip("cv2")
    env = so.SnakeOracle(D, S, S, 3, "classic", draws=np.random.RandomState(1))
    rng = np.random.RandomState(2)
    ob = env.reset()
    r = 84 // (D + 2)
    for _ in range(30):
        warped = cv2.resize(ob, (84, 84), interpolation=cv2.INTER_AREA)
        assert warped.shape == (84, 84, 9)
        assert np.array_equal(warped, np.repeat(np.repeat(ob, r, axis=0), r, axis=1))
        ob, _, done, _ = env.step(rng.randint(0, 5, size=S))
        if done:
            ob = env.reset()


@pytest.mark.reference
def test_reference_warpframe_wrapper_matches():
    """The reference's own WarpFrame class over its own env (build container only)."""
    pytest.importorskip("cv2")
    import ref_loader
    env = ref_loader.make_env("classic", 2, 19, np.random.RandomState(3))
    import importlib
    utils = importlib.import_module("utils")  # /root/reference/src/utils.py
    wrapped = utils.WarpFrame(env)
    ob = wrapped.reset()
    assert ob.shape == (84, 84, 9)
    raw = env.get_multi_snake_ob()
    assert np.array_equal(ob, np.repeat(np.repeat(raw, 4, axis=0), 4, axis=1))
