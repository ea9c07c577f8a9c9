"""Drop-in for the reference's `modules/ar/utils/model.py` (skeleton path).

Same class names, constructor arguments, state_dict schema and `forward`
signature / return dict as the reference `TRXOS` (model.py:219-328); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
`include/arx.h` (libarx.so).  PyTorch is used for parameter storage, device
memory and streams only.  Inference only (the reference calls the model under
`torch.no_grad()`, ar.py:68); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from itertools import combinations

import torch
from torch import nn

from . import _lib


class PositionalEncoding(nn.Module):
    """Buffer holder for `transformers.i.pe.pe` (reference model.py:12-28)."""

    def __init__(self, d_model, dropout, max_len=5000, pe_scale_factor=0.1):
        super().__init__()
        self.pe_scale_factor = pe_scale_factor
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term) * self.pe_scale_factor
        pe[:, 1::2] = torch.cos(position * div_term) * self.pe_scale_factor
        self.register_buffer("pe", pe.unsqueeze(0))


class TemporalCrossTransformer(nn.Module):
    """Parameter holder with the reference's attribute surface (model.py:31-57):
    `k_linear`, `v_linear`, `norm_k`, `pe`, `tuples`, `tuples_len`, `scores`."""

    def __init__(self, args, temporal_set_size=3, add_hook=False):
        super().__init__()
        self.args = args
        self.temporal_set_size = temporal_set_size
        max_len = int(args.seq_len * 1.5)
        self.pe = PositionalEncoding(args.trans_linear_in_dim, args.trans_dropout, max_len=max_len)
        self.k_linear = nn.Linear(args.trans_linear_in_dim * temporal_set_size, args.trans_linear_out_dim)
        self.v_linear = nn.Linear(args.trans_linear_in_dim * temporal_set_size, args.trans_linear_out_dim)
        self.norm_k = nn.LayerNorm(args.trans_linear_out_dim)
        self.tuples_len = math.comb(args.seq_len, temporal_set_size)
        self.add_hook = add_hook
        self.scores = []
        self._tuples = None
        self._owner = None          # set by TRXOS: (model, index) for the device-built table

    @property
    def tuples(self):
        """List of int64 tensors, one per tuple (reference model.py:54).  Fetched from the
        device-built table (arx_tuple_table) when a CUDA handle exists; the host listing is used
        only for this attribute on CPU-only construction, never for scoring."""
        if self._tuples is None:
            tab = None
            if self._owner is not None:
                tab = self._owner[0]()._device_tuple_table(self._owner[1])
            if tab is None:
                tab = torch.tensor(list(combinations(range(self.args.seq_len), self.temporal_set_size)),
                                   dtype=torch.int64)
            self._tuples = [t for t in tab.to(torch.int64)]
        return self._tuples


class MLP(nn.Module):
    """Parameter holder (reference model.py:164-180)."""

    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.input_size, self.hidden_size, self.output_size = input_size, hidden_size, output_size
        self.fc1 = nn.Linear(input_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, output_size)


class Discriminator(nn.Module):
    """Parameter holder (reference model.py:183-204)."""

    def __init__(self, seq_len=120, dim=128, l=16):
        super().__init__()
        self.dimensionality_reduction = nn.Linear(dim, l)
        self.fc1 = nn.Linear(seq_len * l, 256)
        self.fc2 = nn.Linear(256, 64)
        self.fc3 = nn.Linear(64, 1)


class PostResNet(nn.Module):
    """rgb-only layer; kept because its parameters are in every reference state_dict (model.py:207-216)."""

    def __init__(self):
        super().__init__()
        self.l1 = nn.Linear(2048, 256)


class _LazyPrototypes:
    """`out['prototypes']`: list of W tensors (b,1,N,D) (model.py:126,146), materialised on first access."""

    def __init__(self, model, query, way):
        self._model, self._query, self._way, self._val = model, query, way, None

    def _get(self):
        if self._val is None:
            _, protos = self._model.debug_attention(self._query, want_probs=False, want_prototypes=True)
            self._val = [protos[:, c:c + 1] for c in range(self._way)]
            self._query = None
        return self._val

    def __len__(self):
        return self._way

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())


class _HostTicket:
    """Handle of one in-flight `score_host_async` request (keeps the host buffers alive)."""

    def __init__(self, model, ticket, query, logits, is_true):
        self._model, self._ticket, self._query, self._logits, self._is_true = model, ticket, query, logits, is_true

    def result(self):
        m = self._model
        _lib.check(_lib.load().arx_score_host_wait(m._h, self._ticket), m._h, "arx_score_host_wait")
        return self._logits, self._is_true


class TRXOS(nn.Module):
    """B200-native TRX-OS scorer with the reference API (model.py:219-328)."""

    def __init__(self, args, add_hook=False):
        super().__init__()
        if args.input_type != "skeleton":
            raise ValueError("isbfsar_b200 implements the skeleton scoring path only (input_type='skeleton')")
        self.args = args
        self.way = args.way
        self.trans_linear_in_dim = args.trans_linear_in_dim
        self.features_extractor = nn.ModuleDict()
        self.features_extractor["sk"] = MLP(args.n_joints * 3, args.n_joints * 3 * 2, 256)
        self.transformers = nn.ModuleList([TemporalCrossTransformer(args, s, add_hook=add_hook) for s in args.temp_set])
        self.model = args.model
        if self.model == "DISC":
            self.discriminator = Discriminator(seq_len=int(((args.seq_len - 1) * args.seq_len) / 2),
                                               dim=args.trans_linear_out_dim, l=args.seq_len)
        elif self.model == "EXP":
            raise ValueError("model='EXP' (torch.exp head, reference model.py:286-287) is not implemented")
        self.post_resnet = PostResNet()
        self.add_hook = add_hook
        import weakref
        for i, t in enumerate(self.transformers):
            t._owner = (weakref.ref(self), i)
        self._h = None
        self._h_device = None
        self._weights_key = None
        self.max_chunk = 0
        self.force_path = 0
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    # ------------------------------------------------------------------ handle plumbing
    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release(self):
        if getattr(self, "_h", None):
            _lib.load().arx_destroy(self._h)
            self._h = None
            self._weights_key = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _device(self):
        return self.features_extractor["sk"].fc1.weight.device

    def _ensure(self):
        """Create the native handle on the parameters' CUDA device and (re)upload weights when they changed."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("TRXOS parameters are on %s: move the model to a CUDA device (`.cuda()`); "
                               "the scoring path has no CPU fallback" % dev)
        lib = _lib.load()
        if self._h is not None and self._h_device != dev:
            self._release()
        with torch.cuda.device(dev):
            if self._h is None:
                a = self.args
                cfg = _lib.ArxConfig()
                cfg.seq_len, cfg.n_joints = a.seq_len, a.n_joints
                cfg.feat_dim, cfg.out_dim = a.trans_linear_in_dim, a.trans_linear_out_dim
                cfg.n_transformers = len(a.temp_set)
                for i, c in enumerate(a.temp_set):
                    cfg.cardinality[i] = c
                cfg.has_discriminator = 1 if self.model == "DISC" else 0
                cfg.max_chunk, cfg.force_path = int(self.max_chunk), int(self.force_path)
                h = C.c_void_p()
                _lib.check(lib.arx_create(C.byref(cfg), C.byref(h)), None, "arx_create")
                self._h, self._h_device, self._weights_key = h, dev, None
            key = tuple((p.data_ptr(), p._version) for p in self._weight_tensors())
            if key != self._weights_key:
                self._upload_weights()
                self._weights_key = key
        return self._h

    def _weight_tensors(self):
        ts = [self.features_extractor["sk"].fc1.weight, self.features_extractor["sk"].fc1.bias,
              self.features_extractor["sk"].fc2.weight, self.features_extractor["sk"].fc2.bias]
        for t in self.transformers:
            ts += [t.pe.pe, t.k_linear.weight, t.k_linear.bias, t.v_linear.weight, t.v_linear.bias,
                   t.norm_k.weight, t.norm_k.bias]
        if self.model == "DISC":
            d = self.discriminator
            ts += [d.dimensionality_reduction.weight, d.dimensionality_reduction.bias, d.fc1.weight, d.fc1.bias,
                   d.fc2.weight, d.fc2.bias, d.fc3.weight, d.fc3.bias]
        return ts

    def _upload_weights(self):
        def ptr(t):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("weights must be contiguous float32")
            return C.c_void_p(t.data_ptr())
        w = _lib.ArxWeights()
        w.on_device = 1
        m = self.features_extractor["sk"]
        w.fc1_w, w.fc1_b, w.fc2_w, w.fc2_b = ptr(m.fc1.weight), ptr(m.fc1.bias), ptr(m.fc2.weight), ptr(m.fc2.bias)
        for i, t in enumerate(self.transformers):
            w.pe[i], w.k_w[i], w.k_b[i] = ptr(t.pe.pe), ptr(t.k_linear.weight), ptr(t.k_linear.bias)
            w.v_w[i], w.v_b[i] = ptr(t.v_linear.weight), ptr(t.v_linear.bias)
            w.ln_g[i], w.ln_b[i] = ptr(t.norm_k.weight), ptr(t.norm_k.bias)
        if self.model == "DISC":
            d = self.discriminator
            w.dr_w, w.dr_b = ptr(d.dimensionality_reduction.weight), ptr(d.dimensionality_reduction.bias)
            w.d1_w, w.d1_b, w.d2_w, w.d2_b = ptr(d.fc1.weight), ptr(d.fc1.bias), ptr(d.fc2.weight), ptr(d.fc2.bias)
            w.d3_w, w.d3_b = ptr(d.fc3.weight), ptr(d.fc3.bias)
        _lib.check(_lib.load().arx_load_weights(self._h, C.byref(w), self._stream()), self._h, "arx_load_weights")

    def _device_tuple_table(self, ti):
        if self._device().type != "cuda":
            return None
        h = self._ensure()
        n = _lib.load().arx_tuple_count(h, ti)
        out = torch.empty((n, self.args.temp_set[ti]), dtype=torch.int32, device=self._device())
        _lib.check(_lib.load().arx_tuple_table(h, ti, C.c_void_p(out.data_ptr()), self._stream()), h, "arx_tuple_table")
        return out

    def tuple_table(self, ti=0):
        """(N,c) int32 CUDA tensor built by the device kernel (bit-exact with itertools.combinations)."""
        return self._device_tuple_table(ti)

    @staticmethod
    def _f32c(x, dev):
        return x.detach().to(device=dev, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ explicit fast API
    def embed(self, frames):
        """MLP features for (..., 3J) frames -> (..., F)   (reference MLP.forward, model.py:175-180)."""
        h = self._ensure()
        dev = self._device()
        x = self._f32c(frames, dev)
        out = torch.empty(x.shape[:-1] + (self.trans_linear_in_dim,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().arx_embed(h, C.c_void_p(x.data_ptr()), x.numel() // x.shape[-1],
                                             C.c_void_p(out.data_ptr()), self._stream()), h, "arx_embed")
        return out

    def set_support(self, poses=None, features=None):
        """Precompute the support-side tuple K/V operands once (poses (W,T,3J) or features (W,T,F))."""
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        with torch.cuda.device(dev):
            if features is not None:
                f = self._f32c(features, dev)
                assert f.dim() == 3, "features must be (W,T,F)"
                _lib.check(lib.arx_set_support_features(h, C.c_void_p(f.data_ptr()), f.shape[0], self._stream()), h,
                           "arx_set_support_features")
            else:
                p = self._f32c(poses, dev)
                assert p.dim() == 3, "poses must be (W,T,3J)"
                _lib.check(lib.arx_set_support_poses(h, C.c_void_p(p.data_ptr()), p.shape[0], self._stream()), h,
                           "arx_set_support_poses")

    def support_features(self):
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        way = lib.arx_support_way(h)
        out = torch.empty((way, self.args.seq_len, self.trans_linear_in_dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.arx_get_support_features(h, C.c_void_p(out.data_ptr()), self._stream()), h,
                       "arx_get_support_features")
        return out

    def score(self, query, want_chosen=False, out=None):
        """query (B,T,3J) on the model's device -> logits (B,W), is_true (B,1) [None without DISC].
        `out=(logits, is_true)` writes into caller-provided contiguous float32 CUDA tensors."""
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        q = self._f32c(query, dev)
        B = q.shape[0]
        way = lib.arx_support_way(h)
        if way < 1:
            raise RuntimeError("score: support set not set")
        if out is not None:
            logits, is_true = out
            assert logits.shape == (B, way) and logits.is_contiguous() and logits.dtype == torch.float32 and logits.device == dev
            assert is_true is None or (is_true.numel() == B and is_true.is_contiguous() and is_true.dtype == torch.float32)
        else:
            logits = torch.empty((B, way), dtype=torch.float32, device=dev)
            is_true = torch.empty((B, 1), dtype=torch.float32, device=dev) if self.model == "DISC" else None
        chosen = torch.empty((B,), dtype=torch.int32, device=dev) if want_chosen else None
        if B == 0:
            return (logits, is_true, chosen) if want_chosen else (logits, is_true)
        with torch.cuda.device(dev):
            _lib.check(lib.arx_score(h, C.c_void_p(q.data_ptr()), B, C.c_void_p(logits.data_ptr()),
                                     C.c_void_p(is_true.data_ptr()) if is_true is not None else None,
                                     C.c_void_p(chosen.data_ptr()) if chosen is not None else None, self._stream()),
                       h, "arx_score")
        return (logits, is_true, chosen) if want_chosen else (logits, is_true)

    def _host_out(self, B, way, out):
        if out is not None:
            logits, is_true = out
            assert logits.shape == (B, way) and logits.dtype == torch.float32 and logits.is_contiguous()
            return logits, is_true
        logits = torch.empty((B, way), dtype=torch.float32).pin_memory()
        is_true = torch.empty((B, 1), dtype=torch.float32).pin_memory() if self.model == "DISC" else None
        return logits, is_true

    def score_host(self, query_cpu, out=None):
        """End to end from HOST memory (pinned for overlap): query (B,T,3J) CPU -> (logits, is_true) CPU tensors.
        `out=(logits, is_true)` reuses caller-provided (pinned) result tensors."""
        h = self._ensure()
        lib = _lib.load()
        q = query_cpu
        assert q.device.type == "cpu" and q.dtype in (torch.float32, torch.float16) and q.is_contiguous()
        B = q.shape[0]
        way = lib.arx_support_way(h)
        logits, is_true = self._host_out(B, way, out)
        fn = lib.arx_score_host if q.dtype == torch.float32 else lib.arx_score_host_f16      # fp16 rows: half the PCIe bytes
        with torch.cuda.device(self._device()):
            _lib.check(fn(h, C.c_void_p(q.data_ptr()), B, C.c_void_p(logits.data_ptr()),
                          C.c_void_p(is_true.data_ptr()) if is_true is not None else None, None), h, "arx_score_host")
        return logits, is_true

    def score_host_async(self, query_cpu, out=None):
        """Streaming form of score_host: enqueue and return a ticket object; `ticket.result()` blocks until the
        (logits, is_true) host tensors are filled.  Up to two requests are kept in flight by the library, so the
        H2D copy of the next request overlaps the scoring of the current one."""
        h = self._ensure()
        lib = _lib.load()
        q = query_cpu
        assert q.device.type == "cpu" and q.dtype in (torch.float32, torch.float16) and q.is_contiguous()
        B = q.shape[0]
        way = lib.arx_support_way(h)
        logits, is_true = self._host_out(B, way, out)
        t = C.c_int64()
        fn = lib.arx_score_host_submit if q.dtype == torch.float32 else lib.arx_score_host_submit_f16
        with torch.cuda.device(self._device()):
            _lib.check(fn(h, C.c_void_p(q.data_ptr()), B, C.c_void_p(logits.data_ptr()),
                          C.c_void_p(is_true.data_ptr()) if is_true is not None else None, None, C.byref(t)), h, "arx_score_host_submit")
        return _HostTicket(self, int(t.value), q, logits, is_true)

    def score_frames(self, frames):
        """Every sliding window of seq_len consecutive frames of a frame stream (n_frames, 3J) against the current support
        set (arx_score_frames): -> logits (n_frames-T+1, W), is_true (n_frames-T+1, 1) [None without DISC]."""
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        x = self._f32c(frames, dev)
        assert x.dim() == 2 and x.shape[1] == self.args.n_joints * 3, "frames must be (n_frames, 3J)"
        B = max(0, x.shape[0] - self.args.seq_len + 1)
        way = lib.arx_support_way(h)
        if way < 1:
            raise RuntimeError("score_frames: support set not set")
        logits = torch.empty((B, way), dtype=torch.float32, device=dev)
        is_true = torch.empty((B, 1), dtype=torch.float32, device=dev) if self.model == "DISC" else None
        if B == 0:
            return logits, is_true
        with torch.cuda.device(dev):
            _lib.check(lib.arx_score_frames(h, C.c_void_p(x.data_ptr()), x.shape[0], C.c_void_p(logits.data_ptr()),
                                            C.c_void_p(is_true.data_ptr()) if is_true is not None else None, None, self._stream()),
                       h, "arx_score_frames")
        return logits, is_true

    def score_episodes(self, query, poses=None, features=None):
        """Training / evaluation call shape (train.py:110-120, compute_fsos.py:89-98): episode i scores query i
        (b,T,3J) against ITS OWN support classes, poses (b,W,T,3J) or features (b,W,T,F).  One batched pass
        (arx_score_episodes); replaces the current support set.  -> logits (b,W), is_true (b,1) [None without DISC]."""
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        q = self._f32c(query, dev)
        sup = self._f32c(features if features is not None else poses, dev)
        assert sup.dim() == 4 and sup.shape[0] == q.shape[0], "support must be (b,W,T,.) with the query's batch size"
        b, way = q.shape[0], sup.shape[1]
        logits = torch.empty((b, way), dtype=torch.float32, device=dev)
        is_true = torch.empty((b, 1), dtype=torch.float32, device=dev) if self.model == "DISC" else None
        if b == 0:
            return logits, is_true
        with torch.cuda.device(dev):
            _lib.check(lib.arx_score_episodes(h, C.c_void_p(sup.data_ptr()), 1 if features is not None else 0, way,
                                              C.c_void_p(q.data_ptr()), b, C.c_void_p(logits.data_ptr()),
                                              C.c_void_p(is_true.data_ptr()) if is_true is not None else None, None,
                                              self._stream()), h, "arx_score_episodes")
        return logits, is_true

    def stream_push(self, frame):
        """One camera frame (3J floats, host) through the resident streaming scorer (arx_stream_push): returns
        (probs (way,) float32 numpy, is_true float, valid bool).  The sliding window lives on the device."""
        import numpy as np
        # per-frame fast path: the weights are re-validated (data_ptr / version of every parameter) whenever the support
        # set is (re)processed and on every batch call, not on every camera frame
        h = self._h if self._h is not None and self._weights_key is not None else self._ensure()
        lib = _lib.load()
        x = np.ascontiguousarray(np.asarray(frame, dtype=np.float32).reshape(-1))
        if x.size != self.args.n_joints * 3:
            raise ValueError("stream_push: frame must have 3*n_joints values")
        way = lib.arx_support_way(h)
        out = np.empty((way + 1,), dtype=np.float32)
        valid = C.c_int32(0)
        rc = lib.arx_stream_push(h, x.ctypes.data, out.ctypes.data, C.byref(valid))       # the handle carries its device
        if rc:
            _lib.check(rc, h, "arx_stream_push")
        return out[:way], out[way:], bool(valid.value)

    def stream_reset(self):
        _lib.check(_lib.load().arx_stream_reset(self._ensure()), self._h, "arx_stream_reset")

    def score_features(self, ti, qfeats):
        """`transformers[ti](support, labels, queries)['logits']` from frame features (B,T,F)."""
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        f = self._f32c(qfeats, dev)
        B = f.shape[0]
        logits = torch.empty((B, lib.arx_support_way(h)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.arx_score_features(h, ti, C.c_void_p(f.data_ptr()), B, C.c_void_p(logits.data_ptr()),
                                              self._stream()), h, "arx_score_features")
        return logits

    def debug_attention(self, query, want_probs=True, want_prototypes=True):
        """Softmax scores P (B,W,N,N) and prototypes (B,W,N,D) of transformers[0] (model.py:110-111,126)."""
        h = self._ensure()
        dev = self._device()
        lib = _lib.load()
        q = self._f32c(query, dev)
        B, way = q.shape[0], lib.arx_support_way(h)
        N, D = self.transformers[0].tuples_len, self.args.trans_linear_out_dim
        if B * way * N * max(N if want_probs else 0, D) * 4 > (8 << 30):
            raise MemoryError("debug_attention: batch too large for materialised attention outputs")
        probs = torch.empty((B, way, N, N), dtype=torch.float32, device=dev) if want_probs else None
        protos = torch.empty((B, way, N, D), dtype=torch.float32, device=dev) if want_prototypes else None
        with torch.cuda.device(dev):
            _lib.check(lib.arx_debug_attention(h, C.c_void_p(q.data_ptr()), B,
                                               C.c_void_p(probs.data_ptr()) if probs is not None else None,
                                               C.c_void_p(protos.data_ptr()) if protos is not None else None,
                                               self._stream()), h, "arx_debug_attention")
        return probs, protos

    STAGES = ("embed_mlp", "kv_projection", "tuple_build_ln", "cross_attention", "open_set_head")

    def profile(self, on=True):
        """Enable/disable the per-stage CUDA-event timers of arx_score (arx_profile_enable)."""
        _lib.check(_lib.load().arx_profile_enable(self._ensure(), 1 if on else 0), self._h, "arx_profile_enable")

    def profile_read(self, reset=True):
        """-> ({stage: accumulated ms}, chunks) since the last reset."""
        ms = (C.c_double * 5)()
        n = C.c_int64()
        _lib.check(_lib.load().arx_profile_read(self._ensure(), ms, C.byref(n), 1 if reset else 0), self._h,
                   "arx_profile_read")
        return dict(zip(self.STAGES, list(ms))), int(n.value)

    def debug_set(self, key, value):
        _lib.check(_lib.load().arx_debug_set(self._ensure(), int(key), int(value)), self._h, "arx_debug_set")

    def launch_count(self):
        return int(_lib.load().arx_launch_count(self._h)) if self._h else 0

    def last_path(self):
        return int(_lib.load().arx_last_path(self._h)) if self._h else 0

    # ------------------------------------------------------------------ reference API
    @staticmethod
    def _shared_across_batch(t):
        return t.shape[0] == 1 or t.stride(0) == 0

    @torch.no_grad()
    def forward(self, ss_data, ss_labels, query_data, ss_features=None):
        """Reference signature (model.py:291): returns {'logits','is_true','prototypes','support_features'}.

        ss_data {"sk": (b,W,T,3J)} or None when ss_features (b,W,T,F) is given; ss_labels (b,W), only
        row 0 is read (model.py:95) and indexes the class axis; query_data {"sk": (b,T,3J)}.
        When the support set is shared by the whole batch (b==1 or an expanded view) all windows are
        scored in one pass; otherwise every batch row is scored against its own support set."""
        if "rgb" in query_data:
            raise ValueError("rgb/hybrid inputs are outside the skeleton scoring path")
        q = query_data["sk"]
        b = q.shape[0]
        dev = self._device()
        labels = [int(v) for v in ss_labels[0].tolist()]
        src = ss_features if ss_features is not None else ss_data["sk"]
        if src.shape[0] not in (1, b):
            raise RuntimeError(f"support batch {src.shape[0]} does not match query batch {b}")
        lab_t = torch.as_tensor(labels, device=src.device, dtype=torch.long)

        def set_from(row):
            sel = row.index_select(0, lab_t)                      # class order = ss_labels[0] (model.py:95-98)
            if ss_features is not None:
                self.set_support(features=sel)
            else:
                self.set_support(poses=sel)

        if self._shared_across_batch(src):
            set_from(src[0])
            logits, is_true = self.score(q)
            feats_all = self._all_support_features(src, ss_features, dev)
        else:
            # every batch row is its own episode (train.py:110-120): all of them in one batched pass
            sel = src.index_select(1, lab_t)                      # class order = ss_labels[0] (model.py:95-98)
            if ss_features is not None:
                logits, is_true = self.score_episodes(q, features=sel)
            else:
                logits, is_true = self.score_episodes(q, poses=sel)
            feats_all = self._all_support_features(src, ss_features, dev)
        out = {"logits": logits, "support_features": feats_all}
        if is_true is not None:
            out["is_true"] = is_true
        if self._shared_across_batch(src):
            out["prototypes"] = _LazyPrototypes(self, q, len(labels))
            if self.add_hook:
                probs, _ = self.debug_attention(q, want_probs=True, want_prototypes=False)
                for c in range(len(labels)):
                    self.transformers[0].scores.append(probs[:, c:c + 1])
        else:
            out["prototypes"] = None
        return out

    def _all_support_features(self, src, ss_features, dev):
        """'support_features' of the return dict: (b,W,T,F) over ALL classes given (model.py:317,328)."""
        if ss_features is not None:
            return ss_features
        if self._shared_across_batch(src):
            return self.embed(src[:1]).expand(src.shape[0], -1, -1, -1)
        return self.embed(src)

    def distribute_model(self):
        """Reference model.py:360-369 only spreads the rgb extractor; nothing to do on the skeleton path."""
        return None

    # ------------------------------------------------------------------ multi-GPU plumbing
    def export_support(self, out=None):
        """Support-set tuple embeddings as one flat float32 CUDA tensor (for `torch.distributed.broadcast`)."""
        h = self._ensure()
        lib = _lib.load()
        way = lib.arx_support_way(h)
        if way < 1:
            raise RuntimeError("export_support: support set not set")
        n = int(lib.arx_support_blob_bytes(h, way))
        blob = out if out is not None else torch.empty((n // 4,), dtype=torch.float32, device=self._device())
        assert blob.numel() == n // 4 and blob.dtype == torch.float32 and blob.is_contiguous()
        with torch.cuda.device(self._device()):
            _lib.check(lib.arx_export_support(h, C.c_void_p(blob.data_ptr()), self._stream()), h, "arx_export_support")
        return blob

    def support_blob_numel(self, way):
        h = self._ensure()
        return int(_lib.load().arx_support_blob_bytes(h, way)) // 4

    def import_support(self, blob, way):
        h = self._ensure()
        b = self._f32c(blob, self._device())          # no copy for a contiguous float32 CUDA tensor: the caller keeps it alive
        assert b.numel() == self.support_blob_numel(way)
        with torch.cuda.device(self._device()):
            _lib.check(_lib.load().arx_import_support(h, C.c_void_p(b.data_ptr()), way, self._stream()), h,
                       "arx_import_support")
