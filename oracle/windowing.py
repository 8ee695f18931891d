"""Oracle of the clip windowing / stitching of the demo script — TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference cuts a clip with `more_itertools.windowed(frames, FRAME_SLICE_LEN, step=FRAME_SLICE_LEN - OVERLAP)`
and filters the `None` padding (scripts/video_sample.py:361-368); `more-itertools` is an un-vendored, UNPINNED
dependency (requirements.txt:8) that is absent from the build container, so `windowed` is restated here from its
published algorithm (more-itertools 8.x-10.x `more.py: windowed`: a deque of maxlen n, a window is emitted every
`step` appended items, a trailing partial window is padded with `fillvalue`).  The stitching rule is
scripts/video_sample.py:480-487: every window after the first drops its first OVERLAP frames."""
from __future__ import annotations

from collections import deque


def windowed(seq, n, fillvalue=None, step=1):
    """more_itertools.windowed, restated (published behaviour):
    windowed([1,2,3,4,5], 3) -> (1,2,3),(2,3,4),(3,4,5);  windowed([1,2,3], 4) -> (1,2,3,None);
    windowed([1,2,3,4,5,6], 3, fillvalue='!', step=2) -> (1,2,3),(3,4,5),(5,6,'!')."""
    if n < 0:
        raise ValueError("n must be >= 0")
    if n == 0:
        yield ()
        return
    if step < 1:
        raise ValueError("step must be >= 1")
    window = deque(maxlen=n)
    i = n
    for item in seq:
        window.append(item)
        i -= 1
        if not i:
            i = step
            yield tuple(window)
    size = len(window)
    if size == 0:
        return
    if size < n:
        yield tuple(window) + (fillvalue,) * (n - size)
    elif 0 < i < min(step, n):
        window.extend((fillvalue,) * i)
        yield tuple(window)


def script_windows(n_frames, size=10, overlap=3):
    """Frame index lists of scripts/video_sample.py:361-368 (None padding filtered)."""
    return [[f for f in w if f is not None] for w in windowed(range(n_frames), size, step=size - overlap)]


def script_stitch(windows_of_frames, overlap=3):
    """scripts/video_sample.py:480-487: which (window, frame) pairs make the output clip, in order."""
    out = []
    for k, w in enumerate(windows_of_frames):
        out += w if k == 0 else w[overlap:]
    return out
