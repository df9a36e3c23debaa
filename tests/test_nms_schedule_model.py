"""CPU: the scheduling arithmetic of sort_nms_kernel (csrc/yx_postprocess.cu) restated in Python and checked
against plain greedy NMS / sorted(): these are the exactness arguments of the kernel, independent of the GPU --

* the all-ascending bitonic network that never enumerates pairs with a partner >= n,
* the merge passes (merge-path start, odd item size, pairs cut at the end of the list),
* the fixed-length lower bound of the cluster's rank merge,
* survivor batches (groups tested against the kept list, survivors resolved later in one batch, a group that does not fit
  tested again) with the kept list dealt round-robin over the CTAs of a cluster,
* the multiply-only decision of the IoU test outside a 2^-20 band around the threshold.

The GPU tests (tests/test_gpu_postprocess.py) check the kernel itself bit for bit against the reference's goldens."""
import bisect

import numpy as np


# ---------------------------------------------------------------- sort ----------------------------------------------------------
def bitonic_asc(keys, n, max_lg=31):
    """sort_keys_asc: blocks of 2^max_lg sorted keys; pairs whose partner is >= n are not enumerated."""
    lg = 1
    while (1 << (lg - 1)) < n and lg <= max_lg:
        size, half = 1 << lg, 1 << (lg - 1)
        fb, rem = n >> lg, n & (size - 1)
        full = fb << (lg - 1)
        for q in range(full + max(0, rem - half)):
            blk, t = (q >> (lg - 1), q & (half - 1)) if q < full else (fb, (size - rem) + (q - full))
            lo, hi = (blk << lg) + t, (blk << lg) + size - 1 - t
            assert 0 <= lo < hi < n
            if keys[lo] > keys[hi]:
                keys[lo], keys[hi] = keys[hi], keys[lo]
        for lgj in range(lg - 2, -1, -1):
            j = 1 << lgj
            for q in range(((n >> (lgj + 1)) << lgj) + max(0, (n & (2 * j - 1)) - j)):
                lo = ((q & ~(j - 1)) << 1) | (q & (j - 1))
                hi = lo | j
                assert 0 <= lo < hi < n
                if keys[lo] > keys[hi]:
                    keys[lo], keys[hi] = keys[hi], keys[lo]
        lg += 1


def merge_sort(keys, n, threads=512):
    """merge_sort_keys: returns (sorted list, number of passes)."""
    MAX = 1 << 64
    src, dst = keys[:], [None] * n
    bitonic_asc(src, n, 4)
    D = 17 if n <= 4096 else 33
    lgL, passes = 4, 0
    while (1 << lgL) < n:
        L = 1 << lgL
        ipp = (2 * L + D - 1) // D
        items = ((n + 2 * L - 1) >> (lgL + 1)) * ipp
        if n <= 14000:
            assert items <= threads                      # one item per thread for every size two buffers fit
        written = 0
        for it in range(items):
            pr = it // ipp
            d, a0 = (it - pr * ipp) * D, pr << (lgL + 1)
            lenA = min(L, n - a0)
            b0 = a0 + lenA
            lenB = min(L, n - b0)
            cnt = min(D, lenA + lenB - d)
            if cnt <= 0:
                continue
            lo, hi = max(0, d - lenB), min(d, lenA)
            while lo < hi:
                mid = (lo + hi) >> 1
                if src[a0 + mid] < src[b0 + d - 1 - mid]:
                    lo = mid + 1
                else:
                    hi = mid
            i, j = lo, d - lo
            ka = src[a0 + i] if i < lenA else MAX
            kb = src[b0 + j] if j < lenB else MAX
            for t in range(cnt):
                ta = ka < kb
                dst[a0 + d + t] = ka if ta else kb
                i, j = i + (1 if ta else 0), j + (0 if ta else 1)
                nx, ln, base = (i, lenA, a0) if ta else (j, lenB, b0)
                v = src[base + nx] if nx < ln else MAX
                ka, kb = (v, kb) if ta else (ka, v)
                written += 1
        assert written == n
        src, dst = dst, src
        lgL, passes = lgL + 1, passes + 1
    return src, passes


def test_bitonic_network_without_padding_sorts_every_length():
    rng = np.random.default_rng(1)
    for n in list(range(0, 70)) + [127, 128, 129, 511, 513, 1000, 1025]:
        k = rng.permutation(10 * n + 10)[:n].tolist()
        got = k[:]
        bitonic_asc(got, n)
        assert got == sorted(k), n


def test_merge_passes_sort_every_length_and_pass_parity_matches_the_kernel_formula():
    rng = np.random.default_rng(2)
    for n in list(range(0, 70)) + [255, 256, 257, 4096, 4097, 4200, 8192, 8193, 8400, 12000]:
        k = rng.permutation(10 * n + 10)[:n].tolist()
        got, passes = merge_sort(k, n)
        assert got == sorted(k), n
        # merge_passes(): n > 16 ? (32 - clz(n - 1)) - 4 : 0 -- decides which buffer a peer CTA reads the run from
        assert passes == ((n - 1).bit_length() - 4 if n > 16 else 0), n


def test_fixed_length_lower_bound_of_the_rank_merge():
    rng = np.random.default_rng(3)
    for m2 in list(range(0, 70)) + [127, 128, 129, 4199, 4200, 8192]:
        run = sorted(rng.permutation(3 * m2 + 5)[:m2].tolist())
        for k in [-1, 0, 1, 3 * m2 + 10] + rng.integers(0, 3 * m2 + 5, 40).tolist():
            lo, hi = 0, m2
            for _ in range(m2.bit_length()):             # 32 - clz(m2) steps, no early exit
                if lo < hi:
                    mid = (lo + hi) >> 1
                    if run[mid] < k:
                        lo = mid + 1
                    else:
                        hi = mid
            assert lo >= hi and lo == bisect.bisect_left(run, k)


# ---------------------------------------------------------------- greedy NMS ----------------------------------------------------
def _suppresses(a, b, thr):
    w = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
    h = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = w * h
    if inter == 0:
        return False
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter) > thr


def _greedy(boxes, scores, cls, thr, J):
    kept = []
    for i in sorted(range(len(scores)), key=lambda i: (-scores[i], i)):
        if not any(abs(cls[k] - cls[i]) <= J and _suppresses(boxes[k], boxes[i], thr) for k in kept):
            kept.append(i)
    return kept


def _cluster_schedule(boxes, scores, cls, thr, J, R, KC, T=64):
    """The kernel's schedule with groups / batches of T rows: rank merge of R sorted runs, kept list dealt over R CTAs (KC
    entries each in 'shared memory', the rest through the global list), survivor batches with re-test on overflow."""
    n = len(scores)
    keys = [(-scores[i], i) for i in range(n)]
    seg = (n + R - 1) // R
    runs = [sorted(keys[min(n, r * seg):min(n, r * seg + seg)]) for r in range(R)]
    order = [None] * n
    for r in range(R):
        for i, k in enumerate(runs[r]):
            pos = i + sum(bisect.bisect_left(runs[rr], k) for rr in range(R) if rr != r)
            assert order[pos] is None
            order[pos] = k[1]
    sbox, scls = [boxes[i] for i in order], [cls[i] for i in order]
    kept, shared = [None] * n, [[] for _ in range(R)]
    nkept = g0 = nb = cbase = retests = 0
    cbox, ccls, cpos = [None] * T, [None] * T, [None] * T
    while g0 < n or nb > 0:
        nk, flush = nkept, True
        if g0 < n:
            gn = min(T, n - g0)
            dead = [t >= gn for t in range(T)]
            for r in range(R):                                       # every CTA tests the group against its share
                own = (nk + R - 1 - r) // R
                for t in range(gn):
                    me, mc = sbox[g0 + t], scls[g0 + t]
                    hit = any(abs(c - mc) <= J and _suppresses(b, me, thr) for b, c in shared[r][:min(own, KC)])
                    hit = hit or any(abs(scls[kept[k * R + r]] - mc) <= J and _suppresses(sbox[kept[k * R + r]], me, thr)
                                     for k in range(KC, own))
                    dead[t] = dead[t] or hit
            surv = dead.count(False)
            if nb + surv <= T:
                if nb == 0:
                    cbase = g0
                for t in range(T):
                    if not dead[t]:
                        cbox[nb], ccls[nb], cpos[nb] = sbox[g0 + t], scls[g0 + t], g0 + t - cbase
                        nb += 1
                g0 += T
                flush = nb >= T - T // 4 or g0 >= n
            else:
                retests += 1
        if not flush or nb == 0:
            continue
        alive, ck = [True] * nb, []
        for i in range(nb):                                           # mask + scan of the batch
            if alive[i]:
                ck.append(i)
                for j in range(i + 1, nb):
                    if abs(ccls[j] - ccls[i]) <= J and _suppresses(cbox[i], cbox[j], thr):
                        alive[j] = False
        for t, i in enumerate(ck):
            e = nk + t
            kept[e] = cbase + cpos[i]
            if e // R < KC:
                assert len(shared[e % R]) == e // R
                shared[e % R].append((cbox[i], ccls[i]))
        nkept, nb = nk + len(ck), 0
    return [order[kept[k]] for k in range(nkept)], retests


def test_survivor_batches_and_dealt_kept_list_equal_greedy_nms():
    rng = np.random.default_rng(4)
    retests = 0
    for trial in range(12):
        n = int(rng.integers(33, 420))
        ctr = rng.uniform(0, 100 if trial % 2 else 40, (n, 2))
        wh = rng.uniform(5, 40, (n, 2))
        boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32).tolist()
        scores = np.round(rng.uniform(0, 1, n), 2).tolist()               # many ties: the anchor index breaks them
        cls = rng.integers(0, 4, n).tolist()
        for J in (0, 1, 10 ** 9):
            want = _greedy(boxes, scores, cls, 0.3, J)
            for R in (1, 2, 8):
                for KC in (1000, 5):
                    got, r = _cluster_schedule(boxes, scores, cls, 0.3, J, R, KC)
                    retests += r
                    assert got == want, (trial, J, R, KC)
    assert retests > 0                                                     # the overflow path was exercised


# ---------------------------------------------------------------- IoU decision --------------------------------------------------
def test_multiply_only_iou_decision_agrees_with_the_ieee_division():
    """suppresses(): inter > rn(rn(thr * uni) * (1 + 2^-20)) => rn(inter / uni) > thr, inter < rn(rn(thr * uni) * (1 - 2^-20)) =>
    not; only the band between takes the division. numpy float32 arithmetic is IEEE round-to-nearest like the kernel's."""
    f = np.float32
    rng = np.random.default_rng(5)
    N = 2_000_000
    up, dn = f(1.00000095367431640625), f(0.99999904632568359375)
    assert float(up) == 1 + 2.0 ** -20 and float(dn) == 1 - 2.0 ** -20
    for thr in (f(0.65), np.nextafter(f(0.65), f(0)), f(0.45), f(0.5), f(0.3), f(0.001), f(0.99)):
        uni = np.exp(rng.uniform(np.log(1e-3), np.log(1e7), N)).astype(f)
        inter = (uni.astype(np.float64) * float(thr) * (1 + rng.normal(0, 3e-6, N))).astype(f)     # ratios around the threshold
        inter = np.nextafter(inter, np.where(rng.random(N) < 0.5, f(0), f(np.inf)).astype(f))
        ref = (inter / uni) > thr
        p = thr * uni
        yes, no = inter > p * up, inter < p * dn
        assert not (yes & ~ref).any() and not (no & ref).any()
        assert 0.05 < (~(yes | no)).mean() < 0.6                           # the band was actually sampled
