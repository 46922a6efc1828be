"""CPU: the video / frame-folder feeder (decode threads -> ring of host buffers), without a GPU.
The decoded frames are checked against a plain sequential cv2 read and, through PIL's LANCZOS (the checker),
against the frames of the chinchess golden fixture, which were minted from the same video."""
import os

import numpy as np
import pytest

import sfv_b200
from sfv_b200.feeder import ArraySource, Feeder, FrameDirSource, VideoSource
from conftest import GOLDEN
from oracle import chinchess, ref_shim

VIDEO = ref_shim.video_path()


def drain(feeder, n, hw):
    out = np.zeros((n,) + hw + (3,), np.uint8)
    seen = np.zeros(n, bool)
    for slot in feeder:
        a = slot.first - feeder.lo
        assert not seen[a:a + slot.n].any()
        out[a:a + slot.n] = slot.buf[:slot.n].numpy()
        seen[a:a + slot.n] = True
        feeder.release(slot)
    assert seen.all()
    return out


def test_array_source_ranges_and_ragged_tail():
    rng = np.random.default_rng(0)
    fr = rng.integers(0, 256, (37, 8, 12, 3), dtype=np.uint8)
    src = ArraySource(fr)
    for lo, hi, batch, dec in ((0, 37, 8, 1), (5, 30, 4, 3), (0, 37, 64, 2), (10, 10, 4, 2), (36, 37, 4, 4)):
        f = Feeder(src, lo, hi, batch=batch, slots=3, n_decoders=dec, pin=False)
        got = drain(f, hi - lo, (8, 12))
        assert np.array_equal(got, fr[lo:hi]), (lo, hi, batch, dec)
        assert f.stats()["frames"] == hi - lo
    assert src.key(7) == "0000000007.jpg"
    with pytest.raises(ValueError):
        Feeder(src, 3, 99, pin=False)


def test_frame_dir_source(tmp_path):
    import cv2
    rng = np.random.default_rng(1)
    fr = rng.integers(0, 256, (5, 16, 24, 3), dtype=np.uint8)
    for i in (3, 0, 4, 1, 2):
        cv2.imwrite(str(tmp_path / f"{i:010d}.png"), cv2.cvtColor(fr[i], cv2.COLOR_RGB2BGR))
    src = FrameDirSource(str(tmp_path))
    assert len(src) == 5 and src.frame_hw == (16, 24) and src.key(2) == "0000000002.png"
    got = drain(Feeder(src, 0, 5, batch=2, n_decoders=2, pin=False), 5, (16, 24))
    assert np.array_equal(got, fr)
    with pytest.raises(FileNotFoundError):
        FrameDirSource(str(tmp_path / "nothing"))


def test_decode_error_reaches_the_consumer():
    class Broken(ArraySource):
        def reader(self, lo, hi):
            def read_into(dst, start, n):
                return 0
            return read_into
    f = Feeder(Broken(np.zeros((4, 8, 8, 3), np.uint8)), 0, 4, batch=2, pin=False)
    with pytest.raises(RuntimeError, match="could not be decoded"):
        for slot in f:
            f.release(slot)


@pytest.mark.skipif(VIDEO is None, reason="reference sample video not present (oracle/_ref/videos)")
def test_video_source_matches_sequential_decode_and_golden_frames():
    import cv2
    from PIL import Image
    src = VideoSource(VIDEO)
    assert len(src) == 480 and src.frame_hw == (432, 768)
    cap = cv2.VideoCapture(VIDEO)
    seq = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        seq.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB))
    seq = np.stack(seq)
    # three decode threads, each seeking to the start of its contiguous sub-range
    f = Feeder(src, 0, 480, batch=32, slots=4, n_decoders=3, pin=False)
    got = drain(f, 480, (432, 768))
    assert np.array_equal(got, seq)
    # a rank's shard in the middle of the video (seek to a non-keyframe)
    got = drain(Feeder(src, 123, 301, batch=16, n_decoders=2, pin=False), 178, (432, 768))
    assert np.array_equal(got, seq[123:301])
    # the golden fixture's frames are these frames through load_img's first LANCZOS pass + the 64-row window
    g = np.load(os.path.join(GOLDEN, "chinchess_480x64x128.npz"))
    gold = chinchess.frames_from_delta(g["frame_delta"])
    H, W = chinchess.HW
    for i in (0, 77, 479):
        img = np.array(Image.fromarray(seq[i]).resize((W, 72), resample=Image.LANCZOS))[:H]
        assert np.array_equal(img, gold[i])
