"""Drop-in for the reference `two_tower_model` inference surface, on B200 kernels.

Mirrors `EmbeddingLayer`, `UserTower`, `AdTower`, `TwoTowerModel` of the reference
(two_tower_model.py:12-314): identical constructor keywords, identical `state_dict()` keys
(`embedding_layer.embeddings.<name>.weight`, `mlp.{0,1,4,5,8}.*`) so
`load_state_dict(checkpoint['model_state_dict'])` from a reference checkpoint works
(inference.py:99-106), identical forward signatures and `.output_dim`.

The parameters are ordinary torch modules (containers only).  The arithmetic of `forward`
is NOT torch: in eval mode on a CUDA device it runs libb2retr.so -
  EmbeddingLayer.forward -> b2r_gather_concat           (bit-exact fp32 row gather + concat)
  UserTower/AdTower.forward -> b2r_tower_forward         (ONE fused kernel: gather -> 3 chained
       tcgen05 GEMMs with BatchNorm folded into W/b, hidden activations kept in shared memory,
       bias+ReLU from TMEM, fp32 L2-normalise epilogue; no host sync per call - see _DeviceFlags)
Training (`.train()` mode, autograd, TwoTowerLoss) is outside the hot-path scope and raises;
so does a CPU tensor: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["EmbeddingLayer", "UserTower", "AdTower", "TwoTowerModel", "fold_tower_weights"]


def _stream(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda_eval(module: nn.Module, t: torch.Tensor, what: str) -> None:
    if module.training:
        raise RuntimeError(f"{what}: only eval-mode inference is implemented on the B200 path "
                           "(call .eval(); training is out of scope)")
    if not t.is_cuda:
        raise RuntimeError(f"{what}: input is on {t.device}; the B200 path needs CUDA tensors "
                           "(there is no CPU fallback)")


class _DeviceFlags:
    """Status word a tower / gather kernel ORs bits into, read WITHOUT a host sync on the hot path.

    The kernels report a bad categorical id (torch: IndexError) and fp16 saturation through one int32 on
    the device.  Reading it with `.item()` would cost a device->host synchronisation per forward - at B = 1
    (`recommend_ads`, inference.py:223-227) that IS the latency.  So every forward enqueues an async copy of
    the word into pinned host memory plus an event; the word is inspected
      * synchronously on the FIRST forward after the weights (re)loaded - a checkpoint whose activations
        leave the fp16 range does so on its first batch, and so does a mis-built id vocabulary;
      * otherwise at the start of the next forward (event already complete: no wait) or in `check()`.
    `sync_checks = True` on the owning module restores a check (and a sync) on every call."""

    def __init__(self, device):
        self.dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = None

    def publish(self):
        """enqueue device -> pinned copy on the current stream"""
        self.host.copy_(self.dev, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(self.dev.device))

    def poll(self, wait: bool) -> int:
        """bits seen so far (0 when nothing is pending or, without `wait`, not yet known); clears them"""
        if self.event is None:
            return 0
        if wait:
            self.event.synchronize()
        elif not self.event.query():
            return 0
        self.event = None
        bits = int(self.host[0])
        if bits:
            self.dev.zero_()
        return bits


class EmbeddingLayer(nn.Module):
    """One `nn.Embedding(card, embedding_dim)` per categorical field, looked up and
    concatenated in dict order (reference two_tower_model.py:12-49)."""

    check_indices = True  # torch raises IndexError on an out-of-range id; so do we (see _DeviceFlags for when)
    sync_checks = None    # None: first forward after a table change is checked synchronously, later ones deferred

    def __init__(self, feature_dims: Dict[str, int], embedding_dim: int = 16):
        super().__init__()
        self.embeddings = nn.ModuleDict({name: nn.Embedding(card, embedding_dim)
                                         for name, card in feature_dims.items()})
        self.embedding_dim = embedding_dim
        self.num_features = len(feature_dims)
        self._ptr_cache = None
        self._flags = None
        self._fresh = True

    def check(self) -> None:
        """Wait for the pending gathers and raise what they reported."""
        if self._flags is not None and self._flags.poll(wait=True) & _lib.TOWER_BAD_INDEX:
            raise IndexError("index out of range in self")

    def _tables(self, device):
        """(device ptr array, device cards array, keep-alive) for the current weights."""
        ws = [emb.weight for emb in self.embeddings.values()]
        sig = tuple((w.data_ptr(), tuple(w.shape)) for w in ws) + (str(device),)
        if self._ptr_cache is None or self._ptr_cache[0] != sig:
            for w in ws:
                if w.dtype != torch.float32 or not w.is_contiguous() or w.device != device:
                    raise RuntimeError("embedding tables must be contiguous fp32 tensors on the input's device")
            ptrs = torch.tensor([w.data_ptr() for w in ws], dtype=torch.int64, device=device)
            cards = torch.tensor([w.shape[0] for w in ws], dtype=torch.int64, device=device)
            self._ptr_cache = (sig, ptrs, cards)
            self._flags = _DeviceFlags(device)
            self._fresh = True
        return self._ptr_cache[1], self._ptr_cache[2]

    def forward(self, categorical_features: torch.Tensor) -> torch.Tensor:
        _need_cuda_eval(self, categorical_features, "EmbeddingLayer.forward")
        lib = _lib.load()
        dev = categorical_features.device
        cat = categorical_features.to(torch.int64).contiguous()
        B, F = cat.shape
        if F < self.num_features:
            raise IndexError(f"expected at least {self.num_features} categorical columns, got {F}")
        if F != self.num_features:
            cat = cat[:, :self.num_features].contiguous()
        ptrs, cards = self._tables(dev)
        flags = self._flags
        if self.check_indices and flags.poll(wait=False) & _lib.TOWER_BAD_INDEX:
            raise IndexError("index out of range in self (reported by an earlier forward)")
        out = torch.empty((B, self.num_features * self.embedding_dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.b2r_gather_concat(ptrs.data_ptr(), cards.data_ptr(), self.num_features,
                                             self.embedding_dim, cat.data_ptr(), B, out.data_ptr(),
                                             out.shape[1], flags.dev.data_ptr(), _stream(dev)))
            if self.check_indices:
                flags.publish()
                if self.sync_checks or (self.sync_checks is None and self._fresh):
                    self._fresh = False
                    self.check()
        return out


def fold_tower_weights(mlp: nn.Sequential):
    """Fold eval-mode BatchNorm1d into the preceding Linear (float64 on the host):
       W' = W * g / sqrt(var + eps),  b' = (b - mean) * g / sqrt(var + eps) + beta.
    Returns [(W fp32 [out,in], b fp32 [out]), ...] for every Linear in order."""
    mods = list(mlp)
    folded = []
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            W = m.weight.detach().double().cpu()
            b = (m.bias.detach().double().cpu() if m.bias is not None
                 else torch.zeros(W.shape[0], dtype=torch.float64))
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
                bn = mods[i + 1]
                g = bn.weight.detach().double().cpu() if bn.affine else torch.ones_like(b)
                beta = bn.bias.detach().double().cpu() if bn.affine else torch.zeros_like(b)
                scale = g / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
                W = W * scale[:, None]
                b = (b - bn.running_mean.detach().double().cpu()) * scale + beta
            folded.append((W.float().contiguous().numpy(), b.float().contiguous().numpy()))
        i += 1
    return folded


class _Tower(nn.Module):
    """Shared body of UserTower / AdTower (reference two_tower_model.py:52-184)."""

    check_indices = True
    sync_checks = None    # see _DeviceFlags: None = synchronous check on the first forward after (re)loading weights
    operand_dtype = None  # None: fp16 operands, bf16 once a saturation is seen; "fp16" / "bf16" pin the format
    force_path = 0        # tests: 1 = layer-by-layer kernels, 2 = fused kernel
    pair = None           # fused kernel on CTA pairs (tcgen05 cta_group::2): None = library default, 0 / 1 force

    def _build(self, feature_dims: Dict[str, int], numerical_dim: int, embedding_dim: int,
               hidden_dims: List[int], output_dim: int, dropout: float) -> None:
        self.embedding_layer = EmbeddingLayer(feature_dims, embedding_dim)
        width = len(feature_dims) * embedding_dim + numerical_dim
        stack: List[nn.Module] = []
        for h in hidden_dims:
            stack += [nn.Linear(width, h), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(dropout)]
            width = h
        stack.append(nn.Linear(width, output_dim))
        self.mlp = nn.Sequential(*stack)
        self.output_dim = output_dim
        self._numerical_dim = numerical_dim
        self._handle = None
        self._handle_sig = None
        self._keep = None
        self._flags = None
        self._fresh = True
        self._ws = None
        self._tensors = None

    # -- native handle, rebuilt whenever a parameter/buffer changes ---------------------
    def _signature(self, device):
        """Cheap change detector (runs on every forward): the in-place version counters of every parameter
        and buffer.  Storage moves (.to / .cuda / load_state_dict(assign=True)) go through `_apply` /
        `_load_from_state_dict`, which drop the handle explicitly."""
        if self._tensors is None:
            self._tensors = list(self.parameters()) + list(self.buffers())
        return (device, tuple(t._version for t in self._tensors))

    def _native(self, device):
        sig = self._signature(device)
        if self._handle is not None and self._handle_sig == sig:
            return self._handle
        self._free()
        lib = _lib.load()
        folded = fold_tower_weights(self.mlp)
        tables = [e.weight for e in self.embedding_layer.embeddings.values()]
        for w in tables:
            if w.device != device or w.dtype != torch.float32 or not w.is_contiguous():
                raise RuntimeError("move the tower to the input's CUDA device first (.to(device))")
        F = len(tables)
        L = len(folded)     # len(hidden_dims) + 1: any depth (reference two_tower_model.py:83-95 loops over the list)
        cards = (C.c_int64 * F)(*[w.shape[0] for w in tables])
        ptrs = (C.c_void_p * F)(*[w.data_ptr() for w in tables])
        tw = _lib.TowerLayers(
            num_fields=F, emb_dim=self.embedding_layer.embedding_dim, num_numerical=self._numerical_dim,
            num_layers=L, cards=cards, tables=ptrs,
            widths=(C.c_int * L)(*[w.shape[0] for w, _ in folded]),
            w=(C.c_void_p * L)(*[w.ctypes.data for w, _ in folded]),
            b=(C.c_void_p * L)(*[b.ctypes.data for _, b in folded]))
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.b2r_tower_create_layers(C.byref(h), C.byref(tw), device.index or 0))
            if self.operand_dtype is not None:
                _lib.check(lib.b2r_tower_set_param(h, b"operand_dtype", {"fp16": 0.0, "bf16": 1.0}[self.operand_dtype]))
            if self.force_path:
                _lib.check(lib.b2r_tower_set_param(h, b"force_path", float(self.force_path)))
            if self.pair is not None:
                _lib.check(lib.b2r_tower_set_param(h, b"pair", float(self.pair)))
        self._handle, self._handle_sig = h, sig
        self._flags = _DeviceFlags(device)
        self._fresh = True
        return h

    @property
    def native_operand_dtype(self) -> str:
        """'fp16' or 'bf16': what the tensor cores are fed right now (None before the first forward)."""
        if self._handle is None:
            return None
        return "bf16" if _lib.load().b2r_tower_get_param(self._handle, b"operand_dtype") == 1.0 else "fp16"

    def _handle_bits(self, bits: int, deferred: bool) -> bool:
        """React to the status bits of an earlier (deferred) or the current forward.  Returns True when the
        caller should rerun the current batch (operands switched to bf16)."""
        if bits & _lib.TOWER_BAD_INDEX and self.check_indices:
            raise IndexError("index out of range in self" + (" (reported by an earlier forward)" if deferred else ""))
        if bits & _lib.TOWER_SATURATED and self.operand_dtype != "fp16":
            import warnings
            _lib.check(_lib.load().b2r_tower_set_param(self._handle, b"operand_dtype", 1.0))
            warnings.warn(f"{type(self).__name__}: an fp16 tensor-core operand exceeded +-65504 "
                          + ("in an earlier forward (its output was clipped); " if deferred else "; ")
                          + "switching this tower to bf16 operands")
            return True
        return False

    def check(self) -> None:
        """Wait for the pending forwards and raise / react to what they reported."""
        if self._flags is not None and self._handle is not None:
            self._handle_bits(self._flags.poll(wait=True), deferred=True)

    def _load_from_state_dict(self, *a, **k):
        self._free()
        return super()._load_from_state_dict(*a, **k)

    def _free(self):
        self._tensors = None
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().b2r_tower_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._free()

    def _apply(self, fn, *a, **k):  # .to()/.cuda() move the tables: drop the cached handle
        self._free()
        return super()._apply(fn, *a, **k)

    def _encode(self, cat: torch.Tensor, num) -> torch.Tensor:
        _need_cuda_eval(self, cat, type(self).__name__ + ".forward")
        lib = _lib.load()
        dev = cat.device
        cat = cat.to(torch.int64).contiguous()
        B, F = cat.shape
        nf = self.embedding_layer.num_features
        if F < nf:
            raise IndexError(f"expected at least {nf} categorical columns, got {F}")
        if F != nf:
            cat = cat[:, :nf].contiguous()
        if self._numerical_dim:
            num = num.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(num.shape) != (B, self._numerical_dim):
                raise RuntimeError(f"numerical_features must be [{B}, {self._numerical_dim}], got {tuple(num.shape)}")
        h = self._native(dev)
        out = torch.empty((B, self.output_dim), dtype=torch.float32, device=dev)
        if B == 0:
            return out
        flags = self._flags
        with torch.cuda.device(dev):
            self._handle_bits(flags.poll(wait=False), deferred=True)     # what earlier forwards reported (no wait)
            for attempt in range(2):
                need = int(lib.b2r_tower_workspace(h, B))     # 0 on the fused path
                if need and (self._ws is None or self._ws.numel() < need or self._ws.device != dev):
                    self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
                _lib.check(lib.b2r_tower_forward(h, cat.data_ptr(), num.data_ptr() if self._numerical_dim else None,
                                                 B, out.data_ptr(), flags.dev.data_ptr(),
                                                 self._ws.data_ptr() if need else None, need, _stream(dev)))
                flags.publish()
                if not (self.sync_checks or (self.sync_checks is None and self._fresh)):
                    break
                # first forward on these weights: look at the status now; a saturated batch is rerun in bf16
                if not self._handle_bits(flags.poll(wait=True), deferred=False):
                    break
            self._fresh = False
        return out


class UserTower(_Tower):
    """concat(embeddings, numericals) -> MLP -> L2-normalised user embedding."""

    def __init__(self, user_feature_dims: Dict[str, int], numerical_dim: int, embedding_dim: int = 16,
                 hidden_dims: List[int] = [512, 256], output_dim: int = 256, dropout: float = 0.3):
        super().__init__()
        self._build(user_feature_dims, numerical_dim, embedding_dim, list(hidden_dims), output_dim, dropout)

    def forward(self, categorical_features: torch.Tensor, numerical_features: torch.Tensor) -> torch.Tensor:
        return self._encode(categorical_features, numerical_features)


class AdTower(_Tower):
    """embeddings -> MLP -> L2-normalised ad embedding."""

    def __init__(self, ad_feature_dims: Dict[str, int], embedding_dim: int = 16,
                 hidden_dims: List[int] = [512, 256], output_dim: int = 256, dropout: float = 0.3):
        super().__init__()
        self._build(ad_feature_dims, 0, embedding_dim, list(hidden_dims), output_dim, dropout)

    def forward(self, categorical_features: torch.Tensor) -> torch.Tensor:
        return self._encode(categorical_features, None)


class TwoTowerModel(nn.Module):
    """User tower + ad tower (reference two_tower_model.py:187-314), inference methods only."""

    def __init__(self, user_feature_dims: Dict[str, int], ad_feature_dims: Dict[str, int], numerical_dim: int,
                 embedding_dim: int = 16, hidden_dims: List[int] = [512, 256], output_dim: int = 256,
                 dropout: float = 0.3, temperature: float = 0.07):
        super().__init__()
        self.user_tower = UserTower(user_feature_dims, numerical_dim, embedding_dim, hidden_dims, output_dim, dropout)
        self.ad_tower = AdTower(ad_feature_dims, embedding_dim, hidden_dims, output_dim, dropout)
        self.temperature = temperature
        self.output_dim = output_dim

    def forward(self, user_categorical, user_numerical, ad_categorical) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.user_tower(user_categorical, user_numerical), self.ad_tower(ad_categorical)

    def predict_scores(self, user_categorical, user_numerical, ad_categorical) -> torch.Tensor:
        u, a = self.forward(user_categorical, user_numerical, ad_categorical)
        return (u * a).sum(dim=1)

    def get_user_embeddings(self, user_categorical, user_numerical) -> torch.Tensor:
        return self.user_tower(user_categorical, user_numerical)

    def get_ad_embeddings(self, ad_categorical) -> torch.Tensor:
        return self.ad_tower(ad_categorical)

    def compute_loss(self, *args, **kwargs):
        raise NotImplementedError("training (compute_loss / TwoTowerLoss) is outside the B200 hot-path scope")
