"""Host side of the scoring engine: state dicts in, device pointers across the C ABI, scores out.

PyTorch is used for device memory, streams and (in bench.py) torch.distributed only; every
arithmetic step of the scoring path runs in libdfs_b200.so.  Mirrors, per scorer, the reference's
batch loops (src/predict.py:100-111, src/predict_hybrid.py:52-78) and ``forward`` signatures.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _native as N

T_FRAMES, N_FEATS = N.T_FRAMES, N.N_FEATS


def _np32(v):
    """state-dict value (numpy or torch, any device) -> contiguous fp32 numpy array."""
    if hasattr(v, "detach"):
        v = v.detach().to("cpu").numpy()
    return np.ascontiguousarray(np.asarray(v), dtype=np.float32)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _conv_bn(sd, conv_key, bn_key, keep):
    w, b = _np32(sd[conv_key + ".weight"]), _np32(sd[conv_key + ".bias"])
    keep += [w, b]
    c = N.ConvBn()
    c.weight, c.bias = _fptr(w), _fptr(b)
    if bn_key is not None:
        arrs = [_np32(sd[f"{bn_key}.{k}"]) for k in ("weight", "bias", "running_mean", "running_var")]
        keep += arrs
        c.bn_weight, c.bn_bias, c.bn_mean, c.bn_var = (_fptr(a) for a in arrs)
    return c


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("dfs_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _features_struct(x, device_index=None):
    """torch CUDA fp32 tensor (B, 321, 180) with arbitrary strides -> dfs_features.  ``device_index``: the scorer's device; a
    tensor on another GPU raises (the kernels would otherwise read the weights of one device from a stream of another)."""
    if x.dim() != 3 or x.shape[1] != T_FRAMES or x.shape[2] != N_FEATS:
        raise ValueError(f"expected features of shape (B, {T_FRAMES}, {N_FEATS}), got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("dfs_b200 scoring takes CUDA tensors (no CPU fallback); use score_host() for host buffers")
    if device_index is not None and x.device.index != device_index:
        raise RuntimeError(f"features live on {x.device} but the scorer was created on cuda:{device_index}")
    if x.dtype != _require_cuda().float32:
        x = x.float()
    sn, st, sf = x.stride()
    if min(st, sf) <= 0 or (x.shape[0] > 1 and sn <= 0):
        x = x.contiguous()
        sn, st, sf = x.stride()
    f = N.Features(x.data_ptr(), x.shape[0], sn, st, sf)
    return f, x


def _stream_ptr(torch, device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Scorer:
    KIND = ""

    def __init__(self):
        self._h = C.c_void_p()
        self.device_index = 0
        self._lib = N.load()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.dfs_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        N.check(self._lib.dfs_model_set_option(self._h, key.encode(), int(value)), "dfs_model_set_option")

    def profile(self, n_ids: int = 4, reset: bool = True):
        """(ms per kernel id, launches per kernel id) since the last reset; needs set_option('profile', 1)."""
        ms = (C.c_double * n_ids)()
        cnt = (C.c_int64 * n_ids)()
        N.check(self._lib.dfs_model_profile(self._h, ms, cnt, n_ids, int(reset)), "dfs_model_profile")
        return list(ms), list(cnt)

    def saturation_count(self):
        """(elements at +-65504, non-finite elements) in the fp16 feature image and activation buffers of the LAST pass
        (dfs_model_saturation_count): a debug census of what the saturating fp16 converts clipped."""
        torch = _require_cuda()
        sat, nonfin = C.c_int64(), C.c_int64()
        with torch.cuda.device(self.device_index):
            N.check(self._lib.dfs_model_saturation_count(self._h, C.byref(sat), C.byref(nonfin), _stream_ptr(torch, self._device(torch))),
                    "dfs_model_saturation_count")
        return int(sat.value), int(nonfin.value)

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.dfs_model_workspace_bytes(self._h))

    def _device(self, torch):
        return torch.device("cuda", self.device_index)

    # -- host buffers: H2D / D2H inside (the e2e path of bench.py) --
    def score_host(self, feats, flag: int = 1):
        """feats: pinned torch CPU tensor or numpy array (B,321,180), fp32 (dfs_score_host) or fp16 (dfs_score_host_f16:
        half the PCIe bytes; the engine quantises to fp16 anyway), each utterance one dense block.  Other dtypes are
        converted to fp32 first; CUDA tensors are rejected (use score()).  Returns numpy fp32 (B,)."""
        torch = _require_cuda()
        slab = _host_slab(feats)
        out = _pinned_result(self, slab.n, 1)[0]
        with torch.cuda.device(self.device_index):
            stream = _stream_ptr(torch, self._device(torch))
            if slab.f16:
                N.check(self._lib.dfs_score_host_f16(self._h, C.c_void_p(slab.ptr), slab.n, slab.time_major, int(flag),
                                                     C.c_void_p(out.data_ptr()), stream), "dfs_score_host_f16")
            else:
                f = N.Features(slab.ptr, slab.n, *slab.strides)
                N.check(self._lib.dfs_score_host(self._h, C.byref(f), int(flag), C.c_void_p(out.data_ptr()), stream), "dfs_score_host")
        del slab
        return out.numpy().copy()


def _pinned_result(owner, n, m):
    """m pinned fp32 result vectors of n scores, kept on `owner` between calls: cudaHostAlloc costs ~1 ms per call, more than the
    D2H copy it serves.  Callers hand out copies, never these buffers."""
    import torch
    cache = getattr(owner, "_result_cache", None)
    if cache is None or cache[0].numel() < n or len(cache) < m:
        cache = [torch.empty(max(n, 1), dtype=torch.float32, pin_memory=True) for _ in range(m)]
        owner._result_cache = cache
    return [c[:n] for c in cache[:m]]


class _HostSlab:
    """A validated host feature table: pointer, utterance count, element strides, fp16 flag (+ the object that owns the bytes)."""
    __slots__ = ("ptr", "n", "strides", "f16", "time_major", "keep")


def _host_slab(feats) -> "_HostSlab":
    """numpy array or torch CPU tensor (B,321,180) -> _HostSlab.  fp32 and fp16 are passed through in place (any other dtype
    is converted to fp32: reading float64 / bfloat16 bytes as fp32 would return garbage scores without an error); a CUDA
    tensor raises.  Each utterance must be one dense 321x180 or 180x321 block, utterances back to back."""
    import torch
    if isinstance(feats, torch.Tensor):
        if feats.device.type != "cpu":
            raise RuntimeError(f"score_host takes HOST features (got a tensor on {feats.device}); use score() for device tensors")
        if feats.dtype not in (torch.float32, torch.float16):
            feats = feats.float()
        f16 = feats.dtype == torch.float16
        ptr, shape, strides, keep = feats.data_ptr(), tuple(feats.shape), tuple(feats.stride()), feats
    else:
        a = np.asarray(feats)
        if a.dtype not in (np.float32, np.float16):
            a = a.astype(np.float32)
        f16 = a.dtype == np.float16
        ptr, shape, strides, keep = a.ctypes.data, a.shape, tuple(s // a.itemsize for s in a.strides), a
    if len(shape) != 3 or shape[1] != T_FRAMES or shape[2] != N_FEATS:
        raise ValueError(f"expected features of shape (B, {T_FRAMES}, {N_FEATS}), got {tuple(shape)}")
    per = T_FRAMES * N_FEATS
    if shape[0] > 1 and strides[0] != per:
        raise ValueError("host features: utterances must be dense and back to back")
    if (strides[1], strides[2]) == (N_FEATS, 1):
        time_major = 0
    elif (strides[1], strides[2]) == (1, T_FRAMES):
        time_major = 1                                  # the (B,321,180) view of [B,180,321] rows (src/predict.py:103-105)
    else:
        raise ValueError("host features: each utterance must be one dense 321x180 (or 180x321) block")
    s = _HostSlab()
    s.ptr, s.n, s.strides, s.f16, s.time_major, s.keep = ptr, int(shape[0]), (per, strides[1], strides[2]), f16, time_major, keep
    return s


class ScorerGroup:
    """Several scorers over ONE upload of a host feature table (dfs_group_*): every slab crosses PCIe once and all
    members score it.  The reference reads the table once per model (src/ensemble.py:105-122, src/predict_hybrid.py:142-145).

        group = ScorerGroup([cnn2d, cnn1d, cae])
        s2, s1, mse = group.score_host(pinned_table)        # flags default: sigmoid on, CAE normaliser if it has one
    """

    def __init__(self, scorers, stage_utts: int = 0):
        scorers = list(scorers)
        if not scorers:
            raise ValueError("ScorerGroup needs at least one scorer")
        if len({s.device_index for s in scorers}) != 1:
            raise ValueError("all scorers of a group must live on one device")
        _require_cuda()
        self.scorers = scorers
        self.device_index = scorers[0].device_index
        self._lib = N.load()
        self._h = C.c_void_p()
        handles = (C.c_void_p * len(scorers))(*[s._h.value for s in scorers])
        N.check(self._lib.dfs_group_create(C.byref(self._h), handles, len(scorers), int(stage_utts)), "dfs_group_create")

    @property
    def stage_utts(self) -> int:
        return int(self._lib.dfs_group_stage_utts(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.dfs_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_flags(self):
        return [int(getattr(s, "has_normalizer", True)) for s in self.scorers]

    def score_host(self, feats, flags=None):
        """feats as in ``_Scorer.score_host``; flags[i] = apply_sigmoid / apply_normalizer of member i.
        Returns a list of numpy fp32 (B,) score vectors, one per member."""
        torch = _require_cuda()
        slab = _host_slab(feats)
        flags = self.default_flags() if flags is None else [int(f) for f in flags]
        if len(flags) != len(self.scorers):
            raise ValueError("one flag per scorer")
        m = len(self.scorers)
        outs = _pinned_result(self, slab.n, m)
        optr = (C.c_void_p * m)(*[o.data_ptr() for o in outs])
        fl = (C.c_int * m)(*flags)
        dev = torch.device("cuda", self.device_index)
        with torch.cuda.device(dev):
            stream = _stream_ptr(torch, dev)
            if slab.f16:
                N.check(self._lib.dfs_group_score_host_f16(self._h, C.c_void_p(slab.ptr), slab.n, slab.time_major, fl, optr, stream),
                        "dfs_group_score_host_f16")
            else:
                f = N.Features(slab.ptr, slab.n, *slab.strides)
                N.check(self._lib.dfs_group_score_host(self._h, C.byref(f), fl, optr, stream), "dfs_group_score_host")
        del slab
        return [o.numpy().copy() for o in outs]


def pinned_empty(shape, dtype="float32", write_combined: bool = False):
    """Page-locked host array from dfs_pinned_alloc (cudaHostAlloc against the current device; optionally write-combined).
    Returns a numpy array that frees the allocation when it is garbage-collected."""
    _require_cuda()
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dt.itemsize, 1)
    lib = N.load()
    p = C.c_void_p()
    N.check(lib.dfs_pinned_alloc(C.byref(p), nbytes, int(write_combined)), "dfs_pinned_alloc")

    buf = (C.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, lib.dfs_pinned_free, C.c_void_p(p.value))     # the array (and every view of it) keeps `buf` alive as its base
    return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)


def _check_precision(precision, allowed=("fp16", "fp32")):
    """"fp16" = tensor-core path (fp16 operands, fp32 accumulation); "fp32" = full fp32 on the CUDA cores (rank-exact evaluation);
    "split" (2D-CNN) = tensor cores with every operand carried as fp16 value + fp16 residual, three MMAs per product."""
    if precision not in allowed:
        raise ValueError(f"precision must be one of {allowed}, got {precision!r}")


class Cnn2dScorer(_Scorer):
    """CNN2D (src/model.py:12-42) on the tcgen05 path.  ``precision="fp32"`` selects the full-fp32 CUDA-core kernels
    (csrc/cnn2d_fp32.cu; ~35x slower) for evaluations where the rank order of near-equal scores matters; ``precision="split"``
    keeps the tensor cores and carries every input sample, activation and weight as fp16 value + fp16 rounding residual
    (three MMAs per product into the fp32 accumulator): fp32-class scores at about a third of the fp16 rate."""
    KIND = "cnn2d"

    def __init__(self, state_dict, device: int = 0, max_chunk: int = 0, precision: str = "fp16"):
        super().__init__()
        _check_precision(precision, ("fp16", "fp32", "split"))
        _require_cuda()
        keep = []
        w = N.Cnn2dWeights()
        w.base_channels = int(np.asarray(_np32(state_dict["conv.0.weight"])).shape[0])
        fcw, fcb = _np32(state_dict["classifier.weight"]), _np32(state_dict["classifier.bias"])
        w.in_features = int(fcw.shape[1] // max(4 * w.base_channels, 1))
        for i, (ck, bk) in enumerate((("conv.0", "conv.1"), ("conv.5", "conv.6"), ("conv.10", "conv.11"))):
            w.conv[i] = _conv_bn(state_dict, ck, bk, keep)
        keep += [fcw, fcb]
        w.fc_weight, w.fc_bias = _fptr(fcw), _fptr(fcb)
        self.device_index = int(device)
        N.check(self._lib.dfs_cnn2d_create(C.byref(self._h), int(device), C.byref(w), int(max_chunk)), "dfs_cnn2d_create")
        if precision != "fp16":
            self.set_option("precision", {"fp32": 1, "split": 2}[precision])

    def score(self, x, apply_sigmoid: bool = False, return_embedding: bool = False):
        """x: CUDA fp32 (B,321,180), any strides.  Returns (B,) logits/scores [, (B,23040) embedding]."""
        torch = _require_cuda()
        f, x = _features_struct(x, self.device_index)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        emb = torch.empty((x.shape[0], 128 * N_FEATS), dtype=torch.float32, device=x.device) if return_embedding else None
        with torch.cuda.device(x.device):
            N.check(self._lib.dfs_cnn2d_score(self._h, C.byref(f), C.c_void_p(out.data_ptr()),
                                              C.c_void_p(emb.data_ptr()) if emb is not None else None,
                                              int(bool(apply_sigmoid)), _stream_ptr(torch, x.device)), "dfs_cnn2d_score")
        return (out, emb) if return_embedding else out


class Cnn1dScorer(_Scorer):
    """CNN1D (src/model_cnn1d.py:12-46)."""
    KIND = "cnn1d"

    def __init__(self, state_dict, device: int = 0, max_chunk: int = 0, precision: str = "fp16"):
        super().__init__()
        _require_cuda()
        _check_precision(precision)
        keep = []
        w = N.Cnn1dWeights()
        w0 = _np32(state_dict["conv.0.weight"])
        w.base_channels, w.in_features = int(w0.shape[0]), int(w0.shape[1])
        for i, (ck, bk) in enumerate((("conv.0", "conv.1"), ("conv.4", "conv.5"), ("conv.8", "conv.9"))):
            w.conv[i] = _conv_bn(state_dict, ck, bk, keep)
        fcw, fcb = _np32(state_dict["classifier.weight"]), _np32(state_dict["classifier.bias"])
        keep += [fcw, fcb]
        w.fc_weight, w.fc_bias = _fptr(fcw), _fptr(fcb)
        self.device_index = int(device)
        N.check(self._lib.dfs_cnn1d_create(C.byref(self._h), int(device), C.byref(w), int(max_chunk)), "dfs_cnn1d_create")
        if precision == "fp32":
            self.set_option("precision", 1)

    def score(self, x, apply_sigmoid: bool = False):
        torch = _require_cuda()
        f, x = _features_struct(x, self.device_index)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(self._lib.dfs_cnn1d_score(self._h, C.byref(f), C.c_void_p(out.data_ptr()), int(bool(apply_sigmoid)),
                                              _stream_ptr(torch, x.device)), "dfs_cnn1d_score")
        return out


class DlqScorer(_Scorer):
    """StatsPool detector ``DeepfakeDetector`` (src/dlqueen_model.py:132-173): Conv1d k5/k3/k3 + BN + GELU encoder on the
    conv1d tensor-core template, masked mean+std pooling and the two-layer head."""
    KIND = "dlq"

    def __init__(self, state_dict, device: int = 0, max_chunk: int = 0):
        super().__init__()
        _require_cuda()
        keep = []
        w = N.DlqWeights()
        w0 = _np32(state_dict["enc.net.0.weight"])
        w.hidden, w.in_ch = int(w0.shape[0]), int(w0.shape[1])
        for i, (ck, bk) in enumerate((("enc.net.0", "enc.net.1"), ("enc.net.4", "enc.net.5"), ("enc.net.8", "enc.net.9"))):
            w.conv[i] = _conv_bn(state_dict, ck, bk, keep)
        arrs = [_np32(state_dict[k]) for k in ("head.0.weight", "head.0.bias", "head.3.weight", "head.3.bias")]
        keep += arrs
        w.fc1_weight, w.fc1_bias, w.fc2_weight, w.fc2_bias = (_fptr(a) for a in arrs)
        self.device_index = int(device)
        N.check(self._lib.dfs_dlq_create(C.byref(self._h), int(device), C.byref(w), int(max_chunk)), "dfs_dlq_create")

    def score(self, x, lengths=None, apply_sigmoid: bool = False):
        """x: CUDA fp32 (B,321,180) view (any strides; the reference's (B,180,T) batch is ``x.transpose(1, 2)``); lengths:
        optional (B,) valid frame counts.  Returns (B,) logits / sigmoid scores."""
        torch = _require_cuda()
        f, x = _features_struct(x, self.device_index)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        lp = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths).to(x.device, torch.int32).contiguous()
            if lengths.numel() != x.shape[0]:
                raise ValueError("lengths must have one entry per utterance")
            lp = C.c_void_p(lengths.data_ptr())
        with torch.cuda.device(x.device):
            N.check(self._lib.dfs_dlq_score(self._h, C.byref(f), lp, C.c_void_p(out.data_ptr()), int(bool(apply_sigmoid)),
                                            _stream_ptr(torch, x.device)), "dfs_dlq_score")
        return out

    def score_host(self, feats, flag: int = 1):
        raise RuntimeError("DlqScorer has no host-buffer pipeline; move the batch to the device and call score()")


class CaeScorer(_Scorer):
    """ConvAutoencoder reconstruction-MSE scorer (src/model_cae.py:23-125 + src/predict_hybrid.py:66-78).
    ``mean``/``std`` are the FeatureNormalizer statistics (src/dataset_cae.py:18-52), optional."""
    KIND = "cae"

    def __init__(self, state_dict, mean=None, std=None, device: int = 0, max_chunk: int = 0, precision: str = "fp16"):
        super().__init__()
        _require_cuda()
        _check_precision(precision)
        keep = []
        w = N.CaeWeights()
        w.base_channels = int(_np32(state_dict["encoder.0.weight"]).shape[0])
        for i, (ck, bk) in enumerate((("encoder.0", "encoder.1"), ("encoder.4", "encoder.5"), ("encoder.8", "encoder.9"),
                                      ("encoder.12", "encoder.13"))):
            w.enc[i] = _conv_bn(state_dict, ck, bk, keep)
        for i, (ck, bk) in enumerate((("decoder.0", "decoder.1"), ("decoder.3", "decoder.4"), ("decoder.6", "decoder.7"),
                                      ("decoder.9", None))):
            w.dec[i] = _conv_bn(state_dict, ck, bk, keep)
        self.has_normalizer = mean is not None
        if mean is not None:
            m, s = _np32(mean), _np32(std)
            keep += [m, s]
            w.norm_mean, w.norm_std = _fptr(m), _fptr(s)
        self.device_index = int(device)
        N.check(self._lib.dfs_cae_create(C.byref(self._h), int(device), C.byref(w), int(max_chunk)), "dfs_cae_create")
        if precision == "fp32":
            self.set_option("precision", 1)

    def score(self, x, apply_normalizer: bool | None = None):
        """Per-utterance MSE between the (normalised) input and its reconstruction; recon never hits HBM."""
        torch = _require_cuda()
        if apply_normalizer is None:
            apply_normalizer = self.has_normalizer
        f, x = _features_struct(x, self.device_index)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(self._lib.dfs_cae_score(self._h, C.byref(f), int(bool(apply_normalizer)), C.c_void_p(out.data_ptr()),
                                            _stream_ptr(torch, x.device)), "dfs_cae_score")
        return out

    def forward(self, x):
        """Compat path of ConvAutoencoder.forward: (recon (B,321,180), latent (B,256,20,11)); x already normalised."""
        torch = _require_cuda()
        f, x = _features_struct(x, self.device_index)
        recon = torch.empty((x.shape[0], T_FRAMES, N_FEATS), dtype=torch.float32, device=x.device)
        latent = torch.empty((x.shape[0], 256, 20, 11), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(self._lib.dfs_cae_forward(self._h, C.byref(f), C.c_void_p(recon.data_ptr()), C.c_void_p(latent.data_ptr()),
                                              _stream_ptr(torch, x.device)), "dfs_cae_forward")
        return recon, latent

    LAYER_SHAPES = ((160, 90, 32), (80, 45, 64), (40, 22, 128), (20, 11, 256), (40, 22, 128), (80, 45, 64), (160, 90, 32))

    def debug_layer(self, x, layer: int, impl: int = 0, apply_normalizer: bool | None = None):
        """Activations after layer 0..6 (enc1..enc4, dec1..dec3) as (B,H,W,C) fp32; impl 0 = tcgen05, 1 = CUDA cores."""
        torch = _require_cuda()
        if apply_normalizer is None:
            apply_normalizer = self.has_normalizer
        f, x = _features_struct(x, self.device_index)
        h, w, c = self.LAYER_SHAPES[layer]
        out = torch.empty((x.shape[0], h, w, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(self._lib.dfs_cae_debug_layer(self._h, C.byref(f), int(impl), int(layer), int(bool(apply_normalizer)),
                                                  C.c_void_p(out.data_ptr()), _stream_ptr(torch, x.device)), "dfs_cae_debug_layer")
        return out

    def score_host(self, feats, flag: int | None = None):
        return super().score_host(feats, int(self.has_normalizer if flag is None else flag))


def fill_features(n: int, first_utt: int = 0, seed: int = 1234, std: float = 3.2, device: int = 0):
    """Device-generated synthetic [n,321,180] fp32 maps (csrc/synth.cu)."""
    torch = _require_cuda()
    dev = torch.device("cuda", device)
    out = torch.empty((n, T_FRAMES, N_FEATS), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.load().dfs_fill_features(C.c_void_p(out.data_ptr()), n, first_utt, seed, std, _stream_ptr(torch, dev)),
                "dfs_fill_features")
    return out
