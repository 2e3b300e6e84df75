"""Stage-2 ranker on the B200: drop-in for the reference's
`transformer_ranker.TransformerRanker` in eval mode.

Same constructor arguments (transformer_ranker.py:214-226), same `state_dict()` keys (so
`load_state_dict(checkpoint['model_state_dict'])` of inference.py:137-141 works), same
`forward(user_categorical, ad_categorical, numerical, mask=None) -> {'ctr', 'engagement', 'revenue'}`
(:332-380).  The arithmetic runs in `libb2retr.so` (`b2r_ranker_forward`, csrc/ranker.cu): tcgen05 GEMMs with
16-bit operands and fp32 accumulation, fp32 residual stream / LayerNorm / cross products.  Training
(`compute_loss`, autograd) is out of scope: the module raises in `.train()` mode.  There is no CPU path.

The reference feeds the encoder a sequence of length one (`x.unsqueeze(1)`, :358), so the attention softmax
is over a single key and equals 1 whatever W_q, W_k and `mask` are: attention(x) = W_o (W_v x + b_v) + b_o.
`fold_ranker_weights` collapses that to one matrix per layer in float64; W_q / W_k stay in the state dict
(checkpoint compatibility) and never reach the device."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .two_tower_model import _DeviceFlags, _need_cuda_eval

TASKS = ("ctr", "engagement", "revenue")


def _attention_params(d_model: int) -> nn.Module:
    m = nn.Module()
    for name in ("W_q", "W_k", "W_v", "W_o"):
        setattr(m, name, nn.Linear(d_model, d_model))
    return m


def _encoder_layer_params(d_model: int, d_ff: int) -> nn.Module:
    m = nn.Module()
    m.self_attention = _attention_params(d_model)
    ff = nn.Module()
    ff.fc1 = nn.Linear(d_model, d_ff)
    ff.fc2 = nn.Linear(d_ff, d_model)
    m.feed_forward = ff
    m.norm1 = nn.LayerNorm(d_model)
    m.norm2 = nn.LayerNorm(d_model)
    return m


def _head_params(d_model: int, dropout: float) -> nn.Sequential:
    return nn.Sequential(nn.Linear(d_model, 256), nn.ReLU(), nn.Dropout(dropout), nn.Linear(256, 64), nn.ReLU(),
                         nn.Dropout(dropout), nn.Linear(64, 1))


def fold_ranker_weights(model: "TransformerRanker") -> dict:
    """Host-side (float64) preparation of the arrays `b2r_ranker_weights` wants: attention collapsed to
    W_o W_v / W_o b_v + b_o, the positional encoding of position 0 added to the projection bias, cross
    weights transposed to [out, in], the three heads stacked.  Returns float32 C-contiguous numpy arrays."""
    f64 = lambda t: t.detach().double().cpu()   # noqa: E731
    out = {}
    out["w_proj"] = f64(model.feature_projection.weight)
    out["b_proj"] = f64(model.feature_projection.bias) + f64(model.positional_encoding)[0, 0]
    wa, ba, g1, b1, wf1, bf1, wf2, bf2, g2, b2 = ([] for _ in range(10))
    for layer in model.transformer_layers:
        att = layer.self_attention
        wo, wv = f64(att.W_o.weight), f64(att.W_v.weight)
        wa.append(wo @ wv)
        ba.append(wo @ f64(att.W_v.bias) + f64(att.W_o.bias))
        g1.append(f64(layer.norm1.weight)); b1.append(f64(layer.norm1.bias))
        wf1.append(f64(layer.feed_forward.fc1.weight)); bf1.append(f64(layer.feed_forward.fc1.bias))
        wf2.append(f64(layer.feed_forward.fc2.weight)); bf2.append(f64(layer.feed_forward.fc2.bias))
        g2.append(f64(layer.norm2.weight)); b2.append(f64(layer.norm2.bias))
    d = model.d_model
    stack = lambda xs, shape: (torch.stack(xs) if xs else torch.zeros((0,) + shape, dtype=torch.float64))  # noqa: E731
    out["w_attn"], out["b_attn"] = stack(wa, (d, d)), stack(ba, (d,))
    out["ln1_g"], out["ln1_b"] = stack(g1, (d,)), stack(b1, (d,))
    out["w_fc1"], out["b_fc1"] = stack(wf1, (1, d)), stack(bf1, (1,))
    out["w_fc2"], out["b_fc2"] = stack(wf2, (d, 1)), stack(bf2, (d,))
    out["ln2_g"], out["ln2_b"] = stack(g2, (d,)), stack(b2, (d,))
    fi = model.feature_interaction
    out["w_cross"] = stack([f64(w).t().contiguous() for w in fi.cross_weights], (d, d))   # xl @ W == W^T applied to xl
    out["b_cross"] = stack([f64(b) for b in fi.cross_biases], (d,))
    heads = [model.prediction_heads[t] for t in model.prediction_heads]
    out["w_h1"] = torch.stack([f64(h[0].weight) for h in heads]); out["b_h1"] = torch.stack([f64(h[0].bias) for h in heads])
    out["w_h2"] = torch.stack([f64(h[3].weight) for h in heads]); out["b_h2"] = torch.stack([f64(h[3].bias) for h in heads])
    out["w_h3"] = torch.stack([f64(h[6].weight)[0] for h in heads]); out["b_h3"] = torch.stack([f64(h[6].bias)[0] for h in heads])
    return {k: np.ascontiguousarray(v.float().numpy()) for k, v in out.items()}


class TransformerRanker(nn.Module):
    """Reference transformer_ranker.py:208-380 (constructor signature, parameter tree, forward contract)."""

    check_indices = True
    operand_dtype = None     # None: fp16 operands, bf16 once a saturation is seen; "fp16" / "bf16" pin the format

    def __init__(self, user_feature_dims: Dict[str, int], ad_feature_dims: Dict[str, int], numerical_dim: int,
                 embedding_dim: int = 32, d_model: int = 256, num_heads: int = 8, num_layers: int = 3,
                 d_ff: int = 1024, max_seq_len: int = 50, dropout: float = 0.1, num_objectives: int = 3):
        super().__init__()
        if d_model % num_heads != 0:
            raise AssertionError("d_model must be divisible by num_heads")   # reference: assert at :27
        self.user_embeddings = nn.ModuleDict({n: nn.Embedding(c, embedding_dim) for n, c in user_feature_dims.items()})
        self.ad_embeddings = nn.ModuleDict({n: nn.Embedding(c, embedding_dim) for n, c in ad_feature_dims.items()})
        width = (len(user_feature_dims) + len(ad_feature_dims)) * embedding_dim + numerical_dim
        self.feature_projection = nn.Linear(width, d_model)
        self.positional_encoding = nn.Parameter(torch.randn(1, max_seq_len, d_model))
        self.transformer_layers = nn.ModuleList([_encoder_layer_params(d_model, d_ff) for _ in range(num_layers)])
        fi = nn.Module()
        fi.cross_weights = nn.ParameterList([nn.Parameter(torch.randn(d_model, d_model)) for _ in range(3)])
        fi.cross_biases = nn.ParameterList([nn.Parameter(torch.randn(d_model)) for _ in range(3)])
        self.feature_interaction = fi
        self.prediction_heads = nn.ModuleDict({t: _head_params(d_model, dropout) for t in TASKS})
        self.d_model = d_model
        self.embedding_dim = embedding_dim
        self.numerical_dim = numerical_dim
        self.d_ff = d_ff
        self._handle = None
        self._sig = None
        self._flags = None
        self._ws = None
        self._tensors = None
        self._fresh = True
        self._graphs = None
        self._graphs_for = None

    # -- native handle, rebuilt whenever a parameter changes (same scheme as the towers) ---------------
    def _signature(self, device):
        if self._tensors is None:
            self._tensors = list(self.parameters()) + list(self.buffers())
        return (device, tuple(t._version for t in self._tensors))

    def _native(self, device):
        sig = self._signature(device)
        if self._handle is not None and self._sig == sig:
            return self._handle
        self._free()
        lib = _lib.load()
        tables = [e.weight for e in self.user_embeddings.values()] + [e.weight for e in self.ad_embeddings.values()]
        for w in tables:
            if w.device != device or w.dtype != torch.float32 or not w.is_contiguous():
                raise RuntimeError("move the ranker to the input's CUDA device first (.to(device))")
        F = len(tables)
        cards = (C.c_int64 * F)(*[w.shape[0] for w in tables])
        ptrs = (C.c_void_p * F)(*[w.data_ptr() for w in tables])
        a = fold_ranker_weights(self)
        rw = _lib.RankerWeights(
            n_user=len(self.user_embeddings), n_ad=len(self.ad_embeddings), emb_dim=self.embedding_dim,
            num_numerical=self.numerical_dim, d_model=self.d_model, d_ff=self.d_ff,
            n_layers=len(self.transformer_layers), n_cross=len(self.feature_interaction.cross_weights),
            n_tasks=len(self.prediction_heads), head1=a["w_h1"].shape[1], head2=a["w_h2"].shape[1],
            cards=C.cast(cards, C.c_void_p), tables=C.cast(ptrs, C.c_void_p),
            **{k: v.ctypes.data for k, v in a.items()})
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.b2r_ranker_create(C.byref(h), C.byref(rw), device.index or 0))
            if self.operand_dtype is not None:
                _lib.check(lib.b2r_ranker_set_param(h, b"operand_dtype", {"fp16": 0.0, "bf16": 1.0}[self.operand_dtype]))
        self._handle, self._sig = h, sig
        self._flags = _DeviceFlags(device)
        self._fresh = True
        return h

    @property
    def native_operand_dtype(self) -> Optional[str]:
        if self._handle is None:
            return None
        return "bf16" if _lib.load().b2r_ranker_get_param(self._handle, b"operand_dtype") == 1.0 else "fp16"

    def _free(self):
        self._tensors = None
        self._graphs = None
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().b2r_ranker_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:      # interpreter shutdown: torch's own module machinery may already be gone
            pass

    def _apply(self, fn, *a, **k):
        self._free()
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._free()
        return super()._load_from_state_dict(*a, **k)

    def _react(self, bits: int) -> bool:
        """True when the batch must be rerun (operands switched to bf16)."""
        if bits & _lib.TOWER_BAD_INDEX and self.check_indices:
            raise IndexError("index out of range in self")
        if bits & _lib.TOWER_SATURATED and self.operand_dtype != "fp16":
            import warnings
            _lib.check(_lib.load().b2r_ranker_set_param(self._handle, b"operand_dtype", 1.0))
            warnings.warn("TransformerRanker: an fp16 tensor-core operand exceeded +-65504; switching to bf16 operands")
            return True
        return False

    def forward(self, user_categorical: torch.Tensor, ad_categorical: torch.Tensor, numerical: torch.Tensor,
                mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """{'ctr': [B], 'engagement': [B], 'revenue': [B]} raw head outputs, fp32 (reference :372-378).
        `mask` is accepted and has no effect, exactly as in the reference: with one key per query the masked
        softmax is still 1 (:72-75)."""
        _need_cuda_eval(self, user_categorical, "TransformerRanker.forward")
        lib = _lib.load()
        dev = user_categorical.device
        ucat = user_categorical.to(torch.int64).contiguous()
        acat = ad_categorical.to(device=dev, dtype=torch.int64).contiguous()
        num = numerical.to(device=dev, dtype=torch.float32).contiguous()
        B = ucat.shape[0]
        nu, na = len(self.user_embeddings), len(self.ad_embeddings)
        if ucat.shape[1] < nu or acat.shape[1] < na:
            raise IndexError("too few categorical columns for the ranker's embedding tables")
        if ucat.shape[1] != nu:
            ucat = ucat[:, :nu].contiguous()
        if acat.shape[1] != na:
            acat = acat[:, :na].contiguous()
        if acat.shape[0] != B or num.shape[0] != B or num.shape[1] != self.numerical_dim:
            raise RuntimeError("TransformerRanker.forward: inconsistent batch / feature shapes")
        h = self._native(dev)
        T = len(self.prediction_heads)
        if B == 0:
            out = torch.empty((T, 0), dtype=torch.float32, device=dev)
            return {t: out[i] for i, t in enumerate(self.prediction_heads)}
        for attempt in range(2):
            out = self._launch(h, dev, ucat, acat, num, B, T)
            self._flags.publish()
            # the ranker runs once per request on 500 rows: a synchronous status check costs nothing next to
            # the D2H copy of the scores that follows (inference.py:258-260)
            if not self._react(self._flags.poll(wait=True)):
                break
            self._graphs = {}          # operand format changed: the captured launch sequences are stale
        return {t: out[i] for i, t in enumerate(self.prediction_heads)}

    # The forward is ~30 short launches (15 GEMMs + row kernels); at the 500 rows of one user they are launch-bound
    # (0.28 ms eager).  For batches up to _GRAPH_MAX_ROWS the sequence is captured once per batch size into a CUDA
    # graph over static input / output / workspace buffers and replayed (B2R_NO_GRAPHS=1 disables).
    _GRAPH_MAX_ROWS = 4096

    def _launch(self, h, dev, ucat, acat, num, B, T):
        import os
        lib = _lib.load()

        def run(uc, ac, nm, out, ws):
            with torch.cuda.device(dev):
                _lib.check(lib.b2r_ranker_forward(h, uc.data_ptr(), ac.data_ptr(), nm.data_ptr(), B, out.data_ptr(),
                                                  self._flags.dev.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  torch.cuda.current_stream(dev).cuda_stream))

        need = int(lib.b2r_ranker_workspace(h, B))
        if B <= self._GRAPH_MAX_ROWS and not os.environ.get("B2R_NO_GRAPHS"):
            if getattr(self, "_graphs", None) is None or getattr(self, "_graphs_for", None) is not h:
                self._graphs, self._graphs_for = {}, h
            ent = self._graphs.get(B)
            if ent is None:
                ent = {"uc": torch.zeros_like(ucat), "ac": torch.zeros_like(acat), "nm": torch.zeros_like(num),
                       "out": torch.empty((T, B), dtype=torch.float32, device=dev),
                       "ws": torch.empty(max(need, 1), dtype=torch.uint8, device=dev)}
                try:
                    run(ent["uc"], ent["ac"], ent["nm"], ent["out"], ent["ws"])    # eager once: kernel attributes
                    torch.cuda.synchronize(dev)
                    self._flags.dev.zero_()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        run(ent["uc"], ent["ac"], ent["nm"], ent["out"], ent["ws"])
                    ent["graph"] = g
                except Exception as exc:   # not capturable here: stay eager for this batch size
                    import warnings
                    warnings.warn(f"TransformerRanker: CUDA-graph capture failed ({type(exc).__name__}: {exc}); staying eager")
                    torch.cuda.synchronize(dev)
                    ent = False
                self._graphs[B] = ent
            if ent:
                ent["uc"].copy_(ucat, non_blocking=True)
                ent["ac"].copy_(acat, non_blocking=True)
                ent["nm"].copy_(num, non_blocking=True)
                ent["graph"].replay()
                return ent["out"].clone()
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        out = torch.empty((T, B), dtype=torch.float32, device=dev)
        run(ucat, acat, num, out, self._ws)
        return out

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training is out of scope of the B200 inference path (reference :382-420)")
