"""Seeded synthetic head tensors for the five BASELINE.json configurations.

numpy-only and deterministic (PCG64 streams), so the same arrays can be regenerated on the GPU box
from a seed instead of being committed.  Every generator post-processes its draw so that candidate
scores are *separated* (relative gap > GAP) per image and away from the confidence thresholds: the
reference's own sort/top-k tie order is implementation-defined, and a margin of many
ulps keeps the order identical under any correctly-rounded-ish exp/sigmoid implementation.
"""
from __future__ import annotations

import zlib
from typing import List, Sequence, Tuple

import numpy as np

GAP = 2e-5          # minimum relative gap between two candidate scores of one image
THR_MARGIN = 1e-4   # minimum relative distance of a score from a confidence threshold

YOLOV8_SIZES = ((80, 80), (40, 40), (20, 20))
YOLOV8_STRIDES = (8.0, 16.0, 32.0)


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def checksum(arrays: Sequence[np.ndarray]) -> int:
    """CRC of the raw bytes — lets a fixture assert that a regenerated input is bit-identical."""
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1), c)
    return c


def _sigmoid64(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))


def _logit64(p):
    p = np.asarray(p, dtype=np.float64)
    return np.log(p) - np.log1p(-p)


def separate_scores(scores: np.ndarray, thresholds: Sequence[float], lo: float, max_iter: int = 200):
    """Return multiplicative nudges f (close to 1) such that scores*f are pairwise separated by GAP
    (relative) among those >= lo, and at least THR_MARGIN away from each threshold.  scores: 1-D float64."""
    s = scores.astype(np.float64).copy()
    active = np.nonzero(s >= lo)[0]
    for _ in range(max_iter):
        changed = False
        for t in thresholds:
            near = active[np.abs(s[active] - t) < THR_MARGIN * t]
            if near.size:
                s[near] = t * (1.0 + 3.0 * THR_MARGIN)
                changed = True
        order = active[np.argsort(s[active], kind="stable")]
        v = s[order]
        close = np.nonzero((v[1:] - v[:-1]) < GAP * v[1:])[0]
        if close.size:
            # push the upper element of each close pair up by a growing amount
            bump = 1.0 + GAP * (2.0 + (np.arange(close.size) % 7))
            s[order[close + 1]] = np.minimum(s[order[close + 1]] * bump, 1.0 - 1e-4)
            changed = True
        if not changed:
            return s
    raise RuntimeError("separate_scores did not converge")


# ------------------------------------------------------------------------------------------------
# YOLOv8 (C1, C2): levels (B, 4*reg_max+nc, h, w)
# ------------------------------------------------------------------------------------------------
def yolov8_head(seed: int, B: int, nc: int = 80, reg_max: int = 16,
                sizes: Sequence[Tuple[int, int]] = YOLOV8_SIZES, strides: Sequence[float] = YOLOV8_STRIDES,
                conf_list: Sequence[float] = (0.001, 0.25), clustered: bool = False,
                cls_mu: float = -18.19, cls_sigma: float = 4.3155) -> List[np.ndarray]:
    """Box logits N(0,3^2), class logits N(cls_mu, cls_sigma^2) (≈2.5 k candidates/img at conf .001 and
    ≈31 at .25 for the 8400-anchor head).  clustered=True plants ~30 objects per image whose
    neighbouring anchors regress (noisily) the same box with a raised class logit, so NMS suppresses."""
    rng = rng_for(seed)
    no = 4 * reg_max + nc
    levels = []
    for (h, w) in sizes:
        x = rng.standard_normal((B, no, h, w), dtype=np.float32)
        x[:, : 4 * reg_max] *= np.float32(3.0)
        x[:, 4 * reg_max:] *= np.float32(cls_sigma)
        x[:, 4 * reg_max:] += np.float32(cls_mu)
        levels.append(x)
    if clustered:
        _plant_objects(rng, levels, sizes, strides, nc, reg_max)
    _separate_yolov8(levels, nc, reg_max, conf_list)
    return levels


def _plant_objects(rng, levels, sizes, strides, nc, reg_max, n_obj=30, radius=2.0):
    B = levels[0].shape[0]
    img_w = sizes[0][1] * strides[0]
    img_h = sizes[0][0] * strides[0]
    for b in range(B):
        for _ in range(n_obj):
            cls = int(rng.integers(0, nc))
            bw, bh = rng.uniform(40.0, 0.45 * img_w), rng.uniform(40.0, 0.45 * img_h)
            cx, cy = rng.uniform(0.15 * img_w, 0.85 * img_w), rng.uniform(0.15 * img_h, 0.85 * img_h)
            x1, y1, x2, y2 = cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2
            for lvl, ((h, w), s) in enumerate(zip(sizes, strides)):
                gx, gy = cx / s - 0.5, cy / s - 0.5
                ix0, ix1 = int(max(0, np.floor(gx - radius))), int(min(w - 1, np.ceil(gx + radius)))
                iy0, iy1 = int(max(0, np.floor(gy - radius))), int(min(h - 1, np.ceil(gy + radius)))
                for iy in range(iy0, iy1 + 1):
                    for ix in range(ix0, ix1 + 1):
                        ax, ay = ix + 0.5, iy + 0.5
                        d = np.array([ax - x1 / s, ay - y1 / s, x2 / s - ax, y2 / s - ay])
                        if np.any(d < 0.2) or np.any(d > reg_max - 1.2):
                            continue
                        d = d + rng.normal(0.0, 0.08, size=4)
                        for side in range(4):
                            t = float(np.clip(d[side], 0.01, reg_max - 1.01))
                            k0 = int(np.floor(t))
                            fr = t - k0
                            col = rng.standard_normal(reg_max).astype(np.float32)
                            col[k0] = np.float32(9.0 + np.log(max(1.0 - fr, 1e-3)))
                            col[k0 + 1] = np.float32(9.0 + np.log(max(fr, 1e-3)))
                            levels[lvl][b, side * reg_max:(side + 1) * reg_max, iy, ix] = col
                        levels[lvl][b, 4 * reg_max + cls, iy, ix] = np.float32(rng.normal(1.5, 1.2))


def _separate_yolov8(levels, nc, reg_max, conf_list):
    B = levels[0].shape[0]
    lo = 0.5 * min(conf_list)
    for b in range(B):
        cls_views = [l[b, 4 * reg_max:].reshape(nc, -1) for l in levels]   # views into the level arrays
        top = np.concatenate([v.max(axis=0) for v in cls_views]).astype(np.float64)
        arg = np.concatenate([v.argmax(axis=0) for v in cls_views])
        s = _sigmoid64(top)
        s2 = separate_scores(s, conf_list, lo)
        moved = np.nonzero(s2 != s)[0]
        off = 0
        for v in cls_views:
            n = v.shape[1]
            m = moved[(moved >= off) & (moved < off + n)]
            for a in m:
                v[arg[a], a - off] = np.float32(_logit64(s2[a]))
            # keep the runner-up class clearly below the winner for every candidate anchor
            off += n
    # second pass: verify in float32 terms (scores recomputed from the stored float32 logits)
    for b in range(B):
        cls_views = [l[b, 4 * reg_max:].reshape(nc, -1) for l in levels]
        top = np.concatenate([v.max(axis=0) for v in cls_views]).astype(np.float64)
        s = _sigmoid64(top)
        c = np.sort(s[s >= lo])
        if c.size > 1 and np.min((c[1:] - c[:-1]) / c[1:]) < 0.25 * GAP:
            raise RuntimeError("yolov8_head: float32 rounding re-created a near-tie; change the seed")


def yolov8_pred(seed: int, B: int, A: int, nc: int = 80, n_clusters: int = 40, per_cluster: int = 12,
                background: int = 1500, img: float = 640.0, conf_lo: float = 0.001,
                conf_list: Sequence[float] = (0.001, 0.25), nm: int = 0) -> np.ndarray:
    """A decoded prediction tensor (B, 4+nc+nm, A) for stage-B (filter + NMS) tests: clustered,
    overlapping boxes so that suppression really happens, scores separated."""
    rng = rng_for(seed)
    pred = np.zeros((B, 4 + nc + nm, A), dtype=np.float32)
    for b in range(B):
        # background anchors: tiny scores everywhere
        pred[b, 4:4 + nc] = (rng.random((nc, A), dtype=np.float32) * np.float32(0.4 * conf_lo))
        pred[b, 0] = rng.uniform(0, img, A).astype(np.float32)
        pred[b, 1] = rng.uniform(0, img, A).astype(np.float32)
        pred[b, 2] = rng.uniform(8, 200, A).astype(np.float32)
        pred[b, 3] = rng.uniform(8, 200, A).astype(np.float32)
        slots = rng.permutation(A)
        used = 0
        top = np.zeros(A, dtype=np.float64)
        cls_of = np.zeros(A, dtype=np.int64)
        for _ in range(n_clusters):
            c = int(rng.integers(0, nc))
            cx, cy = rng.uniform(60, img - 60, 2)
            w, h = rng.uniform(30, 260, 2)
            k = int(rng.integers(max(1, per_cluster // 2), per_cluster * 2))
            for _ in range(k):
                if used >= A:
                    break
                a = slots[used]
                used += 1
                pred[b, 0, a] = cx + rng.normal(0, 0.04 * w)
                pred[b, 1, a] = cy + rng.normal(0, 0.04 * h)
                pred[b, 2, a] = w * np.exp(rng.normal(0, 0.06))
                pred[b, 3, a] = h * np.exp(rng.normal(0, 0.06))
                top[a] = rng.uniform(0.05, 0.98)
                # sometimes a neighbouring class on the same spot (class-awareness must keep both)
                cls_of[a] = c if rng.random() < 0.85 else int(rng.integers(0, nc))
        nb = min(background, A - used)
        for a in slots[used:used + nb]:
            top[a] = float(np.exp(rng.uniform(np.log(conf_lo * 0.5), np.log(0.6))))
            cls_of[a] = int(rng.integers(0, nc))
        s2 = separate_scores(top, conf_list, 0.5 * conf_lo)
        idx = np.nonzero(top > 0)[0]
        pred[b, 4 + cls_of[idx], idx] = s2[idx].astype(np.float32)
        if nm:
            pred[b, 4 + nc:] = rng.standard_normal((nm, A), dtype=np.float32)
    return pred


# ------------------------------------------------------------------------------------------------
# CenterNet (C3): pred (B, H, W, nc + 4) NHWC = heatmap logits | reg (2) | wh (2)
# ------------------------------------------------------------------------------------------------
def _centernet_peaks64(logits: np.ndarray):
    """Reference-quirk peaks of one image in float64: 3x3 max over (x, class) of the (H, W, nc) logits."""
    pad = np.full((logits.shape[0], logits.shape[1] + 2, logits.shape[2] + 2), -np.inf)
    pad[:, 1:-1, 1:-1] = logits
    m = np.full(logits.shape, -np.inf)
    for dx in range(3):
        for dc in range(3):
            if dx == 1 and dc == 1:
                continue
            m = np.maximum(m, pad[:, dx:dx + logits.shape[1], dc:dc + logits.shape[2]])
    return m  # max over the 8 neighbours


def centernet_pred(seed: int, B: int, H: int = 128, W: int = 128, nc: int = 80, K: int = 100,
                   hm_mu: float = -5.0, hm_sigma: float = 1.5, conf_list: Sequence[float] = (0.001, 0.1)) -> np.ndarray:
    """Heatmap logits N(hm_mu, hm_sigma^2), reg U(0,1), wh U(0,20).  The K+40 best peaks of
    every image get pairwise-separated scores and a clear margin over their neighbourhood."""
    rng = rng_for(seed)
    pred = np.empty((B, H, W, nc + 4), dtype=np.float32)
    pred[..., :nc] = rng.standard_normal((B, H, W, nc), dtype=np.float32) * np.float32(hm_sigma) + np.float32(hm_mu)
    pred[..., nc:nc + 2] = rng.random((B, H, W, 2), dtype=np.float32)
    pred[..., nc + 2:] = rng.random((B, H, W, 2), dtype=np.float32) * np.float32(20.0)
    top = K + 40
    for b in range(B):
        for _ in range(60):
            lg = pred[b, ..., :nc].astype(np.float64)
            nmax = _centernet_peaks64(lg)
            peak = lg >= nmax
            flat = np.where(peak.reshape(-1), lg.reshape(-1), -np.inf)
            idx = np.argpartition(-flat, top)[:top]
            idx = idx[np.argsort(-flat[idx], kind="stable")]
            s = _sigmoid64(flat[idx])
            s2 = separate_scores(s, conf_list, 0.0)
            margin = flat[idx] - nmax.reshape(-1)[idx]
            changed = False
            hm = pred[b, ..., :nc]
            for j in np.nonzero(s2 != s)[0]:
                hm[np.unravel_index(idx[j], hm.shape)] = np.float32(_logit64(s2[j]))
                changed = True
            for j in np.nonzero(margin < 1e-3)[0]:            # lift a peak clearly above its neighbourhood
                hm[np.unravel_index(idx[j], hm.shape)] += np.float32(0.01)
                changed = True
            if not changed:
                break
        else:
            raise RuntimeError("centernet_pred: separation did not converge; change the seed")
    return pred


# ------------------------------------------------------------------------------------------------
# SSD (C4): loc (B, P, 4), conf logits (B, P, nc + 1) with column 0 = background
# ------------------------------------------------------------------------------------------------
def ssd_head(seed: int, B: int, P: int = 8732, nc: int = 20, conf_list: Sequence[float] = (0.001, 0.7),
             boost_frac: float = 0.003) -> Tuple[np.ndarray, np.ndarray]:
    """loc N(0,1); background logit N(12,1), foreground N(0,2^2), and `boost_frac` of the priors get one
    random class raised to background + N(2,2^2) (~2 k pairs > .001 and ~19 > .7 per image).
    Candidate probabilities are separated per image."""
    rng = rng_for(seed)
    loc = rng.standard_normal((B, P, 4), dtype=np.float32)
    conf = rng.standard_normal((B, P, nc + 1), dtype=np.float32) * np.float32(2.0)
    conf[..., 0] = rng.standard_normal((B, P), dtype=np.float32) + np.float32(12.0)
    n_boost = max(1, int(P * boost_frac))
    for b in range(B):
        pri = rng.choice(P, n_boost, replace=False)
        cls = rng.integers(1, nc + 1, n_boost)
        conf[b, pri, cls] = conf[b, pri, 0] + rng.normal(2.0, 2.0, n_boost).astype(np.float32)
    lo = 0.5 * min(conf_list)
    for b in range(B):
        for _ in range(100):
            x = conf[b].astype(np.float64)
            e = np.exp(x - x.max(axis=1, keepdims=True))
            p = e / e.sum(axis=1, keepdims=True)
            fg = p[:, 1:]
            idx = np.nonzero(fg.reshape(-1) >= lo)[0]
            s = fg.reshape(-1)[idx]
            s2 = separate_scores(s, conf_list, lo)
            moved = np.nonzero(s2 != s)[0]
            if moved.size == 0:
                break
            for j in moved:
                pr, c = divmod(int(idx[j]), nc)
                # logit shift that moves the probability from s to s2 (other logits fixed)
                conf[b, pr, c + 1] += np.float32(np.log(s2[j] * (1 - s[j]) / (s[j] * (1 - s2[j]))))
        else:
            raise RuntimeError("ssd_head: separation did not converge; change the seed")
    return loc, conf


# ------------------------------------------------------------------------------------------------
# YOLOv7 (C5): 3 levels (B, 3*(5+nc), s, s), s in {20, 40, 80}
# ------------------------------------------------------------------------------------------------
YOLOV7_SIZES = ((20, 20), (40, 40), (80, 80))


def yolov7_head(seed: int, B: int, nc: int = 80, sizes: Sequence[Tuple[int, int]] = YOLOV7_SIZES,
                conf_list: Sequence[float] = (0.001, 0.5), obj_mu: float = -9.0, obj_sigma: float = 3.0,
                cls_mu: float = -1.0, cls_sigma: float = 2.0) -> List[np.ndarray]:
    """xywh logits N(0,1), objectness N(obj_mu, obj_sigma^2), class logits N(cls_mu, cls_sigma^2)
    (~6 k candidates/img >= .001 and ~36 >= .5).  Candidate scores obj*max_cls are separated
    per image by nudging the objectness logit."""
    rng = rng_for(seed)
    attrs = 5 + nc
    levels = []
    for (h, w) in sizes:
        x = rng.standard_normal((B, 3, attrs, h, w), dtype=np.float32)
        x[:, :, 4] = x[:, :, 4] * np.float32(obj_sigma) + np.float32(obj_mu)
        x[:, :, 5:] = x[:, :, 5:] * np.float32(cls_sigma) + np.float32(cls_mu)
        levels.append(x.reshape(B, 3 * attrs, h, w))
    lo = 0.5 * min(conf_list)
    for b in range(B):
        views = [l[b].reshape(3, attrs, -1) for l in levels]      # views into the level arrays
        for _ in range(100):
            obj = np.concatenate([_sigmoid64(v[:, 4]).reshape(-1) for v in views])
            cls = np.concatenate([_sigmoid64(v[:, 5:].max(axis=1)).reshape(-1) for v in views])
            s = obj * cls
            s2 = separate_scores(s, conf_list, lo)
            moved = np.nonzero(s2 != s)[0]
            if moved.size == 0:
                break
            off = 0
            for v in views:
                n = v.shape[0] * v.shape[2]
                for i in moved[(moved >= off) & (moved < off + n)]:
                    a, cell = divmod(int(i - off), v.shape[2])
                    target_obj = min(s2[i] / cls[i], 1.0 - 1e-6)
                    v[a, 4, cell] = np.float32(_logit64(target_obj))
                off += n
        else:
            raise RuntimeError("yolov7_head: separation did not converge; change the seed")
    return levels


# ------------------------------------------------------------------------------------------------
# YOLOv3 (VOC): 3 levels (B, 3*(5+nc), s, s), s in {13, 26, 52} for a 416^2 input
# ------------------------------------------------------------------------------------------------
YOLOV3_SIZES = ((13, 13), (26, 26), (52, 52))


def yolov3_head(seed: int, B: int, nc: int = 20, sizes: Sequence[Tuple[int, int]] = YOLOV3_SIZES,
                conf_list: Sequence[float] = (0.001, 0.6), obj_mu: float = -10.0, obj_sigma: float = 2.5,
                cls_mu: float = -3.0, cls_sigma: float = 2.0, boost_frac: float = 0.004,
                merged: bool = False) -> List[np.ndarray]:
    """xy logits N(0,1), wh logits N(0,0.5^2), objectness N(obj_mu, obj_sigma^2), class logits
    N(cls_mu, cls_sigma^2); `boost_frac` of the anchors become objects (objectness N(3,1.5^2), one class
    N(2,2^2)), neighbouring cells of a level included so that NMS has overlaps to suppress.  Scores sigmoid(obj)*sigmoid(cls_c) of all (anchor, class) pairs are separated
    per image - or over the whole batch with merged=True (the reference Decoder flattens the batch) - by
    nudging the class logit of the pair."""
    rng = rng_for(seed)
    attrs = 5 + nc
    levels = []
    for (h, w) in sizes:
        x = rng.standard_normal((B, 3, attrs, h, w), dtype=np.float32)
        x[:, :, 2:4] *= np.float32(0.5)
        x[:, :, 4] = x[:, :, 4] * np.float32(obj_sigma) + np.float32(obj_mu)
        x[:, :, 5:] = x[:, :, 5:] * np.float32(cls_sigma) + np.float32(cls_mu)
        n_boost = max(1, int(B * 3 * h * w * boost_frac))
        bb, aa = rng.integers(0, B, n_boost), rng.integers(0, 3, n_boost)
        yy, xx, cc = rng.integers(0, h, n_boost), rng.integers(0, w, n_boost), rng.integers(0, nc, n_boost)
        for dx in (0, 1):                                               # the cell and its right neighbour
            xs = np.minimum(xx + dx, w - 1)
            x[bb, aa, 4, yy, xs] = rng.normal(3.0, 1.5, n_boost).astype(np.float32)
            x[bb, aa, 5 + cc, yy, xs] = rng.normal(2.0, 2.0, n_boost).astype(np.float32)
        levels.append(x.reshape(B, 3 * attrs, h, w))
    lo = 0.5 * min(conf_list)
    groups = [list(range(B))] if merged else [[b] for b in range(B)]
    for grp in groups:
        views = [l[b].reshape(3, attrs, -1) for b in grp for l in levels]   # views into the level arrays
        for _ in range(100):
            obj = [_sigmoid64(v[:, 4]) for v in views]                        # (3, cells)
            sc = [o[:, None, :] * _sigmoid64(v[:, 5:]) for o, v in zip(obj, views)]   # (3, nc, cells)
            flat = np.concatenate([s.reshape(-1) for s in sc])
            idx = np.nonzero(flat >= lo)[0]
            s = flat[idx]
            s2 = separate_scores(s, conf_list, lo)
            moved = np.nonzero(s2 != s)[0]
            if moved.size == 0:
                break
            bounds = np.cumsum([0] + [x.size for x in sc])
            for j in moved:
                i = int(idx[j])
                vi = int(np.searchsorted(bounds, i, side="right") - 1)
                a, c, cell = np.unravel_index(i - bounds[vi], sc[vi].shape)
                target = min(s2[j] / obj[vi][a, cell], 1.0 - 1e-6)
                views[vi][a, 5 + c, cell] = np.float32(_logit64(target))
        else:
            raise RuntimeError("yolov3_head: separation did not converge; change the seed")
    return levels
