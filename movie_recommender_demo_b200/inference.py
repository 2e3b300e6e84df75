"""Drop-in for the reference `inference.AdRecommenderInference` (inference.py:21-331).

Stage 1 (`recommend_ads`, inference.py:223-235) runs on the B200 kernels: user tower on the
device -> the embedding stays in HBM (no `.cpu().numpy()` round trip, inference.py:229) ->
`FAISSIndex.search(k=stage1_k)`.  Stage 2 (the transformer ranker over *random* ad features,
inference.py:241-263) runs on `transformer_ranker.TransformerRanker` of this package
(tcgen05 GEMMs, csrc/ranker.cu) — `batch_recommend` ranks the candidates of ALL users in one call — or on
whatever module with the same forward contract the caller injects; without one, the stage-1 order is returned.

The constructor keeps the reference signature `(model_dir, device)` and its on-disk layout
(`preprocessor.pkl`, `two_tower_best.pt|two_tower_final.pt`, `faiss_index.bin` (+`.metadata`)).
The reference's own `data_preprocessing.CriteoDataPreprocessor` (CPU ETL — out of scope, unchanged) is
imported by name if present on sys.path; every component can also be injected, which is how the
tests exercise this class without the reference checkout.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from .faiss_retrieval import FAISSIndex, TwoStageRetriever
from .two_tower_model import TwoTowerModel

__all__ = ["AdRecommenderInference"]

USER_CAT_COLS = [f'C{i}' for i in range(1, 7)]
AD_CAT_COLS = [f'C{i}' for i in range(7, 27)]


def _load_checkpoint(model: torch.nn.Module, path: Path, device) -> None:
    ckpt = torch.load(path, map_location=device)
    state = ckpt['model_state_dict'] if isinstance(ckpt, dict) and 'model_state_dict' in ckpt else ckpt
    model.load_state_dict(state)


class AdRecommenderInference:
    def __init__(self, model_dir: str = '/home/claude/ad_recommender/models',
                 device: str = 'cuda' if torch.cuda.is_available() else 'cpu', *,
                 preprocessor=None, two_tower_model: Optional[TwoTowerModel] = None,
                 transformer_ranker=None, faiss_index: Optional[FAISSIndex] = None, verbose: bool = True):
        self.model_dir = Path(model_dir)
        self.device = device
        self.verbose = verbose
        if 'cuda' not in str(device):
            raise RuntimeError("AdRecommenderInference on the B200 path needs a CUDA device (no CPU fallback)")
        self._say("Loading models and index...")
        self.preprocessor = preprocessor if preprocessor is not None else self._load_preprocessor()
        dims = self.preprocessor.feature_dims
        self.user_feature_dims = {c: dims[c] for c in USER_CAT_COLS if c in dims}
        self.ad_feature_dims = {c: dims[c] for c in AD_CAT_COLS if c in dims}
        self.numerical_dim = len(self.preprocessor.numerical_cols)
        self.two_tower_model = (two_tower_model if two_tower_model is not None else self._load_two_tower())
        self.two_tower_model = self.two_tower_model.to(device).eval()
        self.transformer_ranker = transformer_ranker if transformer_ranker is not None else self._load_transformer()
        if self.transformer_ranker is not None:
            self.transformer_ranker = self.transformer_ranker.to(device).eval()
        self.faiss_index = faiss_index if faiss_index is not None else self._load_faiss_index()
        self.retriever = TwoStageRetriever(self.two_tower_model, self.transformer_ranker, self.faiss_index,
                                           device=self.device)
        self._encoders = None
        self._say("✓ Models loaded successfully!")

    def _say(self, msg: str) -> None:
        if self.verbose:
            print(msg)

    # ------------------------------------------------------------------ loading (reference layout)
    def _load_preprocessor(self):
        try:
            from data_preprocessing import CriteoDataPreprocessor  # the reference's module, unchanged
        except ImportError as exc:
            raise ImportError("pass preprocessor=... or put the reference's data_preprocessing.py on sys.path") from exc
        pre = CriteoDataPreprocessor()
        pre.load(str(self.model_dir / 'preprocessor.pkl'))
        return pre

    def _load_two_tower(self) -> TwoTowerModel:
        model = TwoTowerModel(user_feature_dims=self.user_feature_dims, ad_feature_dims=self.ad_feature_dims,
                              numerical_dim=self.numerical_dim, embedding_dim=16, hidden_dims=[512, 256],
                              output_dim=256, dropout=0.3)
        path = self.model_dir / 'two_tower_best.pt'
        if not path.exists():
            path = self.model_dir / 'two_tower_final.pt'
        _load_checkpoint(model, path, self.device)
        return model

    def _load_transformer(self):
        path = self.model_dir / 'transformer_ranker_best.pt'
        if not path.exists():
            path = self.model_dir / 'transformer_ranker_final.pt'
        if not path.exists():
            return None
        from .transformer_ranker import TransformerRanker   # same ctor / state-dict keys as the reference's module
        model = TransformerRanker(user_feature_dims=self.user_feature_dims, ad_feature_dims=self.ad_feature_dims,
                                  numerical_dim=self.numerical_dim, embedding_dim=32, d_model=256, num_heads=8,
                                  num_layers=3, d_ff=1024, dropout=0.1)
        _load_checkpoint(model, path, self.device)
        return model

    def _load_faiss_index(self) -> FAISSIndex:
        index = FAISSIndex(dimension=256, index_type='IVF', nlist=100, nprobe=10, use_gpu=True)
        index.load(str(self.model_dir / 'faiss_index.bin'))
        self._say(f"  FAISS Index: {index.index.ntotal:,} ads indexed")
        return index

    # ------------------------------------------------------------------ preprocessing
    def _encoder_tables(self):
        """value -> code dicts built once from the fitted LabelEncoders (the reference calls
        `encoder.transform([value])` per field per request, inference.py:172-181)."""
        if self._encoders is None:
            self._encoders = {}
            for col in USER_CAT_COLS:
                enc = getattr(self.preprocessor, "label_encoders", {}).get(col)
                if enc is not None:
                    self._encoders[col] = {v: i for i, v in enumerate(enc.classes_)}
        return self._encoders

    def preprocess_user_features(self, user_data: dict) -> tuple:
        cat, num = self.preprocess_user_batch([user_data])
        return cat, num

    def preprocess_user_batch(self, user_data_list: list) -> tuple:
        """Vectorised `preprocess_user_features` (inference.py:160-197) for a list of users."""
        tables = self._encoder_tables()
        cats = []
        for user in user_data_list:
            row = []
            for col in USER_CAT_COLS:
                if col not in tables:
                    continue
                table = tables[col]
                value = user['categorical'].get(col, 'missing')
                if value not in table:
                    if 'missing' not in table:
                        raise ValueError(f"y contains previously unseen labels: {value!r}")  # sklearn's message
                    value = 'missing'
                row.append(table[value])
            cats.append(row)
        user_categorical = torch.tensor(cats, dtype=torch.long)
        raw = np.array([[user['numerical'].get(col, 0) for col in self.preprocessor.numerical_cols]
                        for user in user_data_list], dtype=np.float64)
        logged = np.log1p(np.abs(raw)).astype(np.float32)
        user_numerical = torch.tensor(self.preprocessor.scaler.transform(logged))
        return user_categorical, user_numerical

    # ------------------------------------------------------------------ inference
    def _stage1(self, user_categorical, user_numerical, stage1_k):
        cat_dev = user_categorical.to(self.device)
        num_dev = user_numerical.to(self.device, dtype=torch.float32)
        self._dev_inputs = (user_categorical, user_numerical, cat_dev, num_dev)   # stage 2 reuses the uploads
        with torch.no_grad():
            user_emb = self.two_tower_model.get_user_embeddings(cat_dev, num_dev)
        return self.faiss_index.search(user_emb, k=stage1_k)   # CUDA tensor in, numpy (ids, scores) out

    def _stage2_batch(self, user_categorical, user_numerical, candidate_ids, top_k, return_scores):
        """Stage 2 for U users at once: candidate_ids [U, stage1_k].  ONE ranker call over U * stage1_k rows
        (the reference makes one call per user, inference.py:250-255, :309-316).  Returns per user
        (order [<= top_k], scores dict | None)."""
        U, stage1_k = candidate_ids.shape[0], candidate_ids.shape[1]
        if self.transformer_ranker is None:
            return [(np.arange(min(top_k, stage1_k)), None) for _ in range(U)]
        held = getattr(self, "_dev_inputs", None)
        if held is not None and held[0] is user_categorical and held[1] is user_numerical:
            cat_dev, num_dev = held[2], held[3]           # uploaded by stage 1 of this request
        else:
            cat_dev = user_categorical.to(self.device)
            num_dev = user_numerical.to(self.device, dtype=torch.float32)
        # every user's features repeated once per candidate (the reference's `.repeat(stage1_k, 1)`, :242-243)
        batch_user_cat = cat_dev[:, None, :].expand(U, stage1_k, cat_dev.shape[1]).reshape(U * stage1_k, -1)
        batch_user_num = num_dev[:, None, :].expand(U, stage1_k, num_dev.shape[1]).reshape(U * stage1_k, -1)
        # the reference scores RANDOM ad features here (inference.py:246-248): one draw per user, in user order,
        # from the global CPU generator.  ONE draw of U * stage1_k rows consumes the same random stream as U draws
        # of stage1_k rows (the CPU generator fills element by element; tests/test_host_logic_cpu.py checks it),
        # so a seeded run sees the numbers the reference's loop would see.
        batch_ad_cat = torch.randint(0, 200, (U * stage1_k, 20)).to(self.device)
        tasks = ('ctr', 'engagement', 'revenue')
        with torch.no_grad():
            pred = self.transformer_ranker(batch_user_cat, batch_ad_cat, batch_user_num)
            # one sigmoid launch and ONE device -> host copy for the three heads (the reference: three of each)
            sig_all = torch.sigmoid(torch.stack([pred[t].reshape(-1) for t in tasks])).cpu().numpy()
        sig = {t: sig_all[i].reshape(U, stage1_k) for i, t in enumerate(tasks)}
        out = []
        for u in range(U):
            ctr = sig['ctr'][u]
            order = np.argsort(ctr)[::-1][:top_k]
            scores = None
            if return_scores:
                scores = {t: sig[t][u][order].tolist() for t in tasks}
            out.append((order, scores))
        return out

    def _stage2(self, user_categorical, user_numerical, candidate_ids, top_k, return_scores):
        return self._stage2_batch(user_categorical, user_numerical, np.asarray(candidate_ids)[None, :], top_k,
                                  return_scores)[0]

    def recommend_ads(self, user_data: dict, top_k: int = 10, stage1_k: int = 500,
                      return_scores: bool = True) -> dict:
        user_categorical, user_numerical = self.preprocess_user_features(user_data)
        self._say(f"\n=== Recommending {top_k} ads ===")
        t0 = time.time()
        candidate_ids, stage1_scores = self._stage1(user_categorical, user_numerical, stage1_k)
        stage1_ms = (time.time() - t0) * 1000
        self._say(f"Stage 1: Retrieved {stage1_k} candidates in {stage1_ms:.2f}ms")
        t1 = time.time()
        order, scores = self._stage2(user_categorical, user_numerical, candidate_ids[0], top_k, return_scores)
        stage2_ms = (time.time() - t1) * 1000
        self._say(f"Stage 2: Ranked to top {top_k} in {stage2_ms:.2f}ms")
        self._say(f"Total: {stage1_ms + stage2_ms:.2f}ms")
        out = {'ad_ids': candidate_ids[0][order].tolist(),
               'timing': {'stage1_ms': stage1_ms, 'stage2_ms': stage2_ms, 'total_ms': stage1_ms + stage2_ms},
               'candidate_ids': candidate_ids[0], 'stage1_scores': stage1_scores[0]}
        if return_scores and scores is not None:
            out['scores'] = scores
        return out

    def batch_recommend(self, user_data_list: list, top_k: int = 10, stage1_k: int = 500) -> list:
        """Same result list as the reference's serial loop (inference.py:290-331), but stage 1 is ONE
        batched tower call + ONE batched search for all users."""
        self._say(f"\n=== Batch Recommending for {len(user_data_list)} users ===")
        t0 = time.time()
        cat, num = self.preprocess_user_batch(user_data_list)
        ids, dist = self._stage1(cat, num, stage1_k)
        stage1_ms = (time.time() - t0) * 1000 / max(len(user_data_list), 1)
        t1 = time.time()
        ranked = self._stage2_batch(cat, num, ids, top_k, True) if len(user_data_list) else []
        stage2_ms = (time.time() - t1) * 1000 / max(len(user_data_list), 1)
        results = []
        for i, (order, scores) in enumerate(ranked):
            rec = {'ad_ids': ids[i][order].tolist(),
                   'timing': {'stage1_ms': stage1_ms, 'stage2_ms': stage2_ms, 'total_ms': stage1_ms + stage2_ms},
                   'candidate_ids': ids[i], 'stage1_scores': dist[i]}
            if scores is not None:
                rec['scores'] = scores
            results.append(rec)
        total = time.time() - t0
        self._say(f"\n✓ Batch complete: {total:.2f}s total ({total / max(len(user_data_list), 1) * 1000:.2f}ms per user)")
        return results
