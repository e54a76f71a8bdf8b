"""Device implementations of the metric half of the hot path, with the reference's signatures:

    calculate_eer(scores, labels) -> (eer, threshold)              scripts/evaluation.py:7-39
    confusion_at_threshold(scores, labels, thr) -> (tp,fp,tn,fn,far,frr)   scripts/evaluation.py:42-56
    normalise_01(scores)                                             src/predict_hybrid.py:81-85
    hybrid_blend(sup, cae, alpha)                                    src/predict_hybrid.py:149-151
    ensemble_mean(all_scores)                                        src/ensemble.py:121

Inputs may be lists, numpy arrays, or torch tensors (CPU or CUDA); fp32 arrays are sorted as fp32
keys, everything else as float64, exactly like ``np.array(scores)`` would type them.  Tie order is
stable by original index (DESIGN.md "EER tie contract").
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("dfs_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _stream(torch, dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _scores_to_device(scores, device=None):
    """-> contiguous CUDA tensor of dtype float32 or float64 (numpy's typing of the input)."""
    torch = _torch()
    if isinstance(scores, torch.Tensor):
        t = scores.detach()
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
    else:
        a = np.array(scores)                       # scripts/evaluation.py:8
        if a.dtype != np.float32:
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    dev = device if device is not None else (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    return t.to(dev).contiguous().reshape(-1)


def _labels_to_device(labels, dev):
    torch = _torch()
    if isinstance(labels, torch.Tensor):
        t = labels.detach().to(dev)
        if t.dtype == torch.uint8:      # the kernels read "label != 0": a uint8 tensor is used in place
            return t.contiguous().reshape(-1), t
        if t.dtype == torch.bool:
            return t.contiguous().reshape(-1).view(torch.uint8), t
        return (t != 0).to(torch.uint8).contiguous().reshape(-1), t
    a = np.array(labels)
    return torch.from_numpy(np.ascontiguousarray((a != 0).astype(np.uint8))).to(dev).reshape(-1), a


def eer_details(scores, labels, want_perm=False, want_sorted=False, device=None, method="auto"):
    """dict(eer, threshold, eer_idx, n_bonafide, n_spoof[, perm][, sorted]) -- eer_idx = -1 on the
    single-class early-out (scripts/evaluation.py:18-19).

    method: "sort" = full stable radix sort + sweep (dfs_eer; the only one that can return perm /
    sorted), "select" = MSD radix select of the FAR/FRR crossing (dfs_eer_select; bit-identical
    result, ~5 B/score per key byte instead of a sort), "auto" = select unless perm / sorted is wanted."""
    if method not in ("auto", "sort", "select"):
        raise ValueError("method must be 'auto', 'sort' or 'select'")
    if method == "select" and (want_perm or want_sorted):
        raise ValueError("the select path does not materialise the permutation; use method='sort'")
    use_select = method == "select" or (method == "auto" and not (want_perm or want_sorted))
    torch = _torch()
    s = _scores_to_device(scores, device)
    lab, _ = _labels_to_device(labels, s.device)
    if s.numel() != lab.numel():
        raise ValueError("scores and labels must have the same length")
    n = s.numel()
    if n == 0:
        raise ValueError("calculate_eer needs at least one score")
    res = N.EerResult()
    perm = torch.empty(n, dtype=torch.int32, device=s.device) if want_perm else None
    srt = torch.empty_like(s) if want_sorted else None
    with torch.cuda.device(s.device):
        if use_select:
            N.check(N.load().dfs_eer_select(C.c_void_p(s.data_ptr()), s.element_size(), C.c_void_p(lab.data_ptr()), n, C.byref(res),
                                            _stream(torch, s.device)), "dfs_eer_select")
        else:
            N.check(N.load().dfs_eer(C.c_void_p(s.data_ptr()), s.element_size(), C.c_void_p(lab.data_ptr()), n, C.byref(res),
                                     C.c_void_p(perm.data_ptr()) if perm is not None else None,
                                     C.c_void_p(srt.data_ptr()) if srt is not None else None, _stream(torch, s.device)), "dfs_eer")
    out = dict(eer=float(res.eer), threshold=float(res.threshold), eer_idx=int(res.eer_idx),
               n_bonafide=int(res.n_bonafide), n_spoof=int(res.n_spoof))
    if want_perm:
        out["perm"] = perm
    if want_sorted:
        out["sorted"] = srt
    return out


def calculate_eer(scores, labels, method="auto"):
    d = eer_details(scores, labels, method=method)
    return d["eer"], d["threshold"]


def confusion_at_threshold(scores, labels, threshold):
    torch = _torch()
    s = _scores_to_device(scores)
    lab_u8, lab_raw = _labels_to_device(labels, s.device)
    # the reference counts label == 1 / label == 0 after .astype(int); keep other values out of both classes
    if isinstance(lab_raw, np.ndarray):
        li = lab_raw.astype(int)
        lab_dev = torch.from_numpy(np.where(li == 1, 1, np.where(li == 0, 0, 2)).astype(np.uint8)).to(s.device)
    else:
        li = lab_raw.to(torch.int64)
        lab_dev = torch.where(li == 1, 1, torch.where(li == 0, 0, 2)).to(torch.uint8)
    out4 = (C.c_int64 * 4)()
    with torch.cuda.device(s.device):
        N.check(N.load().dfs_confusion(C.c_void_p(s.data_ptr()), s.element_size(), C.c_void_p(lab_dev.data_ptr()), s.numel(),
                                       float(threshold), out4, _stream(torch, s.device)), "dfs_confusion")
    tp, fp, tn, fn = (int(v) for v in out4)
    far = fp / (fp + tn) if (fp + tn) > 0 else 0.0
    frr = fn / (tp + fn) if (tp + fn) > 0 else 0.0
    return tp, fp, tn, fn, float(far), float(frr)


def _f64_device(x, dev=None):
    torch = _torch()
    if isinstance(x, torch.Tensor):
        t = x.detach()
        d = dev if dev is not None else (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        t = t.to(d)
        if t.dtype == torch.float32:            # fp32 model scores -> float64 column, on the device
            out = torch.empty(t.numel(), dtype=torch.float64, device=d)
            t = t.contiguous().reshape(-1)
            with torch.cuda.device(d):
                N.check(N.load().dfs_widen_f32_f64(C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(out.data_ptr()), _stream(torch, d)),
                        "dfs_widen_f32_f64")
            return out
        return t.double().contiguous().reshape(-1)
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).reshape(-1)
    d = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.from_numpy(a).to(d)


def blend(score_list, weights, minmax_flags, divisor=1.0, as_numpy=True):
    """out = (sum_m weights[m] * (minmax? normalise_01(s_m) : s_m)) / divisor in float64 on the device, summed left to right like
    numpy.  dfs_blend_f64 takes up to 8 vectors per call; longer lists (src/ensemble.py accepts any number of checkpoints) are
    folded eight at a time with the running sum as the first operand of the next call (weight 1, no min-max, divisor 1: exact),
    which keeps the very same left-to-right order."""
    torch = _torch()
    if not (len(score_list) == len(weights) == len(minmax_flags)) or len(score_list) < 1:
        raise ValueError("blend takes one weight and one min-max flag per score vector (at least one)")
    ts = [_f64_device(score_list[0])]
    ts += [_f64_device(s, ts[0].device) for s in score_list[1:]]
    n = ts[0].numel()
    if any(t.numel() != n for t in ts):
        raise ValueError("all score vectors must have the same length")
    weights, flags = [float(v) for v in weights], [int(bool(v)) for v in minmax_flags]
    out = None
    while ts:
        take = 8 if out is None else 7
        part, pw, pf = ts[:take], weights[:take], flags[:take]
        ts, weights, flags = ts[take:], weights[take:], flags[take:]
        if out is not None:
            part, pw, pf = [out] + part, [1.0] + pw, [0] + pf
        nxt = torch.empty(n, dtype=torch.float64, device=part[0].device)
        m = len(part)
        ptrs = (C.c_void_p * m)(*[t.data_ptr() for t in part])
        w = (C.c_double * m)(*pw)
        f = (C.c_int * m)(*pf)
        with torch.cuda.device(nxt.device):
            N.check(N.load().dfs_blend_f64(ptrs, m, w, f, float(divisor) if not ts else 1.0, n, C.c_void_p(nxt.data_ptr()),
                                           _stream(torch, nxt.device)), "dfs_blend_f64")
        out = nxt
    return out.cpu().numpy() if as_numpy else out


def normalise_01(scores, as_numpy=True):
    return blend([scores], [1.0], [1], 1.0, as_numpy)


def hybrid_blend(sup_scores, cae_scores, alpha=0.80, as_numpy=True):
    return blend([sup_scores, cae_scores], [alpha, 1 - alpha], [1, 1], 1.0, as_numpy)


def ensemble_mean(all_scores, as_numpy=True):
    return blend(list(all_scores), [1.0] * len(all_scores), [0] * len(all_scores), float(len(all_scores)), as_numpy)


def alpha_sweep(sup_scores, cae_scores, labels, alphas=None, alpha_steps=21):
    """The alpha sweep of src/hybrid_ensemble.py:131-151 on the device: both score vectors are min-max normalised
    once (float64), then for every alpha the blend ``alpha*sup + (1-alpha)*cae`` and its EER are computed without the
    vectors leaving the GPU (the reference re-sorts on the host 21 times via ``.tolist()``).

    Returns dict(alphas, eer [A], threshold [A], best_alpha, best_eer) -- best = first strict improvement over 1.0,
    like the reference's ``if eer < best_eer`` loop."""
    torch = _torch()
    if alphas is None:
        alphas = np.linspace(0.0, 1.0, int(alpha_steps))                # hybrid_ensemble.py:132
    sup_n = normalise_01(sup_scores, as_numpy=False)                    # :128-129
    cae_n = _f64_device(normalise_01(cae_scores, as_numpy=False), sup_n.device)
    lab, _ = _labels_to_device(labels, sup_n.device)
    eers, thrs = [], []
    best_eer, best_alpha = 1.0, 0.0
    for a in alphas:
        a = float(a)
        combined = blend([sup_n, cae_n], [a, 1 - a], [0, 0], 1.0, as_numpy=False)   # :145, float64
        d = eer_details(combined, lab)
        eers.append(d["eer"])
        thrs.append(d["threshold"])
        if d["eer"] < best_eer:
            best_eer, best_alpha = d["eer"], a
    return dict(alphas=np.asarray(alphas, dtype=np.float64), eer=np.array(eers), threshold=np.array(thrs),
                best_alpha=best_alpha, best_eer=best_eer)


def bce_with_logits_mean(logits, labels):
    """nn.BCEWithLogitsLoss()(logits, labels) over the whole vector in one fused device reduction (fp64 sum of the fp32
    per-element loss) -- the avg_loss of src/evaluation.py::evaluate."""
    torch = _torch()
    x = logits.detach() if isinstance(logits, torch.Tensor) else torch.as_tensor(np.asarray(logits, dtype=np.float32))
    dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = x.to(dev, torch.float32).contiguous().reshape(-1)
    y = labels.detach() if isinstance(labels, torch.Tensor) else torch.as_tensor(np.asarray(labels, dtype=np.float32))
    y = y.to(dev, torch.float32).contiguous().reshape(-1)
    if x.numel() != y.numel():
        raise ValueError("logits and labels must have the same length")
    if x.numel() == 0:
        raise ValueError("bce_with_logits_mean needs at least one score")
    out = C.c_double()
    with torch.cuda.device(dev):
        N.check(N.load().dfs_bce_with_logits(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), x.numel(), C.byref(out), _stream(torch, dev)),
                "dfs_bce_with_logits")
    return float(out.value)
