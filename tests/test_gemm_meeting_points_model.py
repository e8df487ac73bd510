"""A model of the drift-bounding meeting points of gemm_tf32_2cta_kernel (rindow_math_matrix_b200/csrc/gemm_tcgen05.cu,
TcParams::sync_kb / sync_cnt): every `sync_kb` k-blocks the producer thread of each CTA adds one to a device-wide counter
and spins until the counter has reached "everything that arrives up to and including this meeting".  A spin on a global
counter is the one construct in the library that could hang a GPU, so its arithmetic is restated here and run under
adversarial schedules: for every shape no CTA may wait for an arrival that never comes, and the counters must be back at
zero when the kernel ends (they are not re-initialised between launches).

The model follows the kernel line by line: tiles `pair, pair + num_pairs, ...` per CTA pair, two CTAs per pair,
`wave_ctas = 2 * min(num_pairs, num_tiles - wave_first)`, a meeting before k-block kb when `kb > 0 and kb % sync_kb == 0`,
`sync_target += wave_ctas` then arrive then wait; on exit each CTA adds one to the second counter and the last one
clears both.
"""
import random

import pytest


def run_model(num_tiles, num_pairs, num_k_blocks, sync_kb, rng, max_steps=2_000_000):
    pairs = min(num_pairs, num_tiles)                 # launch_2cta_t: pairs = min(tiles, sm_count / 2)
    grid = 2 * pairs
    counter = [0, 0]                                  # sync_cnt[0] arrivals, sync_cnt[1] finished CTAs

    def cta(block):
        pair = block >> 1
        target = 0
        for tile in range(pair, num_tiles, pairs):
            wave_first = tile - pair
            wave_ctas = 2 * min(pairs, num_tiles - wave_first)
            for kb in range(num_k_blocks):
                if sync_kb > 0 and kb > 0 and kb % sync_kb == 0:
                    target += wave_ctas
                    counter[0] += 1                   # atomicAdd
                    while counter[0] < target:
                        yield "spin"
                yield "load"                          # the TMA loads of this k-block
        counter[1] += 1
        if counter[1] == grid:                        # the last CTA to leave resets the counters
            counter[0] = 0
            counter[1] = 0
        yield "exit"

    live = {b: cta(b) for b in range(grid)}
    spinning = set()
    steps = 0
    while live:
        steps += 1
        assert steps < max_steps, "no progress: a CTA waits for an arrival that never comes"
        runnable = [b for b in live if b not in spinning] or list(live)
        b = rng.choice(runnable)
        state = next(live[b], "exit")
        if state == "spin":
            spinning.add(b)
            if len(spinning) == len(live):            # everybody waits: re-poll all; if nobody moves it is a deadlock
                before = counter[0]
                moved = False
                for s in list(spinning):
                    st = next(live[s], "exit")
                    if st != "spin":
                        spinning.discard(s)
                        moved = True
                        if st == "exit":
                            del live[s]
                assert moved or counter[0] != before, "deadlock: every CTA spins at a meeting point"
        else:
            spinning.discard(b)
            if state == "exit":
                del live[b]
        if state != "spin" and spinning:
            spinning.clear()                          # the counter may have moved: let the spinners poll again
    return counter


SHAPES = [
    # (num_tiles, num_pairs, num_k_blocks, sync_kb)
    (4096, 74, 512, 64),      # 16384^3: 64 x 64 pair tiles of 256 x 256, K = 16384
    (1024, 74, 256, 64),      # 8192^3
    (75, 74, 256, 64),        # one tile in the last wave
    (147, 74, 130, 64),       # one tile short of two full waves
    (5, 74, 256, 64),         # fewer tiles than pairs: the grid shrinks
    (149, 74, 65, 64),        # exactly one meeting per tile
    (149, 74, 64, 64),        # no meeting at all (kb never reaches 64)
    (33, 7, 200, 16),
    (32, 8, 129, 1),          # a meeting before every k-block
    (9, 4, 50, 0),            # switched off
]


@pytest.mark.parametrize("shape", SHAPES, ids=["%dt_%dp_%dk_s%d" % s for s in SHAPES])
def test_meeting_points_never_deadlock_and_reset(shape):
    num_tiles, num_pairs, num_k_blocks, sync_kb = shape
    # keep the pure-Python schedule short: scale the K loop of the two production shapes down, meetings kept
    if num_tiles * num_k_blocks > 40_000:
        scale = max(1, (num_tiles * num_k_blocks) // 40_000)
        num_tiles = max(num_pairs + 1, num_tiles // scale)
    for seed in range(3):
        counter = run_model(num_tiles, num_pairs, num_k_blocks, sync_kb, random.Random(seed))
        assert counter == [0, 0]


def test_one_cta_racing_ahead_is_held_at_the_next_meeting():
    """The throttle itself: a CTA that is always scheduled first cannot get more than one meeting interval ahead."""
    num_tiles, pairs, kblocks, sync_kb = 8, 4, 40, 8
    counter = [0]
    progress = {}

    def cta(block):
        target = 0
        for tile in range(block >> 1, num_tiles, pairs):
            wave_ctas = 2 * min(pairs, num_tiles - (tile - (block >> 1)))
            for kb in range(kblocks):
                if kb > 0 and kb % sync_kb == 0:
                    target += wave_ctas
                    counter[0] += 1
                    while counter[0] < target:
                        yield
                progress[block] = (tile // pairs) * kblocks + kb
                yield

    live = {b: cta(b) for b in range(2 * pairs)}
    order = sorted(live)
    while live:
        for b in order:                               # CTA 0 always first: it would run away without the meetings
            if b in live and next(live[b], "done") == "done":
                del live[b]
        if len(progress) == 2 * pairs and live:
            assert max(progress.values()) - min(progress.values()) <= sync_kb
