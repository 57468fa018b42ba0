"""CPU restatements of three pieces of kernel logic whose correctness is an argument, not arithmetic -- pinned here so the
argument is executable (the GPU parity tests cover the kernels themselves):

* composite_fwd4.cu `Blend`: the running-product blend state (T multiplied unconditionally, `final_T` trailing) takes the
  same decisions and produces the same numbers, bit for bit, as App. A.3's formulation with an explicit "done" flag;
* composite_fwd4.cu / composite_bwd4.cu: the circular survivor queue never overwrites an unread survivor and never hands
  out a group that straddles the wrap;
* binning.cu `merge_chunks_kernel`: the branch-free 11-step search + the "all 2048 below" fix-up is a lower bound.
"""
import numpy as np

f32 = np.float32
ALPHA_MIN, ALPHA_MAX, T_MIN = f32(1.0 / 255.0), f32(0.99), f32(0.0001)


def _blend_reference(alpha_raw, power_ok, col):
    """App. A.3: skip unless power <= 0 and alpha >= 1/255; a splat that would take T below 1e-4 ends the pixel unapplied."""
    T, C, last, done = f32(1), f32(0), 0, False
    for k, (ar, ok, c) in enumerate(zip(alpha_raw, power_ok, col)):
        if done or not ok or ar < ALPHA_MIN:
            continue
        al = min(ALPHA_MAX, ar)
        test = f32(T * f32(f32(1) - al))
        if test < T_MIN:
            done = True
            continue
        C = f32(C + f32(c * f32(al * T)))
        T, last = test, k + 1
    return T, C, last


def _blend_running_product(alpha_raw, power_ok, col):
    """composite_fwd4.cu blend_group4: T is multiplied for every accepted splat, applied or not; Tf trails it."""
    T, Tf, C, last = f32(1), f32(1), f32(0), 0
    for k, (ar, pk, c) in enumerate(zip(alpha_raw, power_ok, col)):
        al = min(ALPHA_MAX, ar)
        ok = bool(ar >= ALPHA_MIN) and bool(pk)
        om = f32(f32(1) - al) if ok else f32(1)
        w = f32(al * T)
        Tn = f32(T * om)
        use = ok and Tn >= T_MIN
        if use:
            C = f32(C + f32(c * w))
            last, Tf = k + 1, Tn
        T = Tn
    return Tf, C, last


def test_running_product_blend_equals_the_flagged_formulation_bit_for_bit():
    rng = np.random.default_rng(0)
    for trial in range(400):
        n = int(rng.integers(1, 400))
        scale = [0.02, 0.2, 1.0, 3.0][trial % 4]  # from "never saturates" to "saturates within a few splats"
        ar = (rng.random(n) ** 2 * scale).astype(f32)
        ar[rng.random(n) < 0.2] = f32(0.001)         # below 1/255: skipped
        pk = rng.random(n) > 0.05                     # power > 0: skipped
        col = rng.random(n).astype(f32)
        a = _blend_reference(ar, pk, col)
        b = _blend_running_product(ar, pk, col)
        assert a[2] == b[2]
        assert np.array([a[0], a[1]], dtype=f32).tobytes() == np.array([b[0], b[1]], dtype=f32).tobytes()


def test_circular_survivor_queue_never_clobbers_and_never_straddles():
    rng = np.random.default_rng(1)
    for group in (4, 8):
        cq = 32 + group
        slots = [None] * cq
        head = fill = 0
        produced = consumed = 0
        for chunk in range(3000):
            hits = int(rng.integers(0, 33)) if chunk % 7 else 32
            for j in range(hits):  # append: slot = head + fill + rank, wrapped once
                slot = head + fill + j
                slot -= cq if slot >= cq else 0
                assert 0 <= slot < cq and slots[slot] is None, "an unread survivor would be overwritten"
                slots[slot] = produced
                produced += 1
            fill += hits
            while fill >= group:  # whole groups, oldest first
                assert head % group == 0 and head + group <= cq, "a group straddles the wrap"
                for u in range(group):
                    assert slots[head + u] == consumed, "survivors leave the queue in list order"
                    slots[head + u] = None
                    consumed += 1
                head = 0 if head + group == cq else head + group
                fill -= group
            assert fill < group
        assert produced - consumed == fill


def _lower_bound_branch_free(sorted_keys, key, chunk=2048):
    """merge_chunks_kernel: number of entries of a sorted chunk (len <= 2048) below `key`."""
    n, l, step = len(sorted_keys), 0, chunk // 2
    while step >= 1:
        probe = l + step
        if probe <= n and sorted_keys[probe - 1] < key:
            l = probe
        step >>= 1
    if n == chunk and l == chunk - 1 and sorted_keys[chunk - 1] < key:
        l = chunk
    return l


def test_branch_free_rank_search_is_a_lower_bound():
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 1000, 2047, 2048):
        keys = np.unique(rng.integers(0, 1 << 40, size=4 * n, dtype=np.int64))[:n]
        assert len(keys) == n
        probes = np.concatenate([keys - 1, keys + 1, [keys[0] - 5, keys[-1] + 5], rng.integers(0, 1 << 40, size=200)])
        for k in probes:  # the merged lists never share a key (depth bits + unique id), so k is never in `keys`
            if k in keys:
                continue
            assert _lower_bound_branch_free(keys, k) == int(np.searchsorted(keys, k, side="left"))
