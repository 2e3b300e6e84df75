"""Generates tests/golden/preprocess_users.npz: the reference's own `AdRecommenderInference.preprocess_user_features`
(inference.py:159-197) on a handful of raw user dicts - unseen category values, absent keys, negative numericals -
over sklearn LabelEncoders / a StandardScaler fitted here (what the reference's CriteoDataPreprocessor holds).
Stored: the fitted parameters (so the test rebuilds the same preprocessor without sklearn), the users, and the
reference's tensors.  Run in the build container only.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def users():
    out = []
    for u in range(12):
        cat = {f"C{i + 1}": f"v{(5 * u + 3 * i) % 9}" for i in range(6)}
        num = {f"I{i + 1}": float((u * 7 + i * 3) % 23) - 4.0 for i in range(13)}      # some negatives: |x| is logged
        if u % 4 == 1:
            cat["C2"] = "never-seen"            # -> 'missing' (inference.py:178-180)
        if u % 4 == 2:
            del cat["C5"]                       # absent key -> 'missing' (:174)
            del num["I3"]                       # absent key -> 0 (:188)
        out.append({"categorical": cat, "numerical": num})
    return out


def main():
    from sklearn.preprocessing import LabelEncoder, StandardScaler
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))     # inference.py imports faiss_retrieval -> faiss
    sys.path.insert(0, "/root/reference")
    import inference as ref                     # the reference, untouched

    rng = np.random.default_rng(11)
    pre = types.SimpleNamespace(label_encoders={}, numerical_cols=[f"I{i + 1}" for i in range(13)])
    for i in range(6):
        enc = LabelEncoder()
        enc.fit([f"v{j}" for j in range(7 + i)] + ["missing"])
        pre.label_encoders[f"C{i + 1}"] = enc
    pre.scaler = StandardScaler().fit(np.log1p(np.abs(rng.normal(3.0, 6.0, (500, 13)))).astype(np.float32))
    inst = object.__new__(ref.AdRecommenderInference)       # no model loading: the method only reads .preprocessor
    inst.preprocessor = pre
    cats, nums = [], []
    for u in users():
        c, n = inst.preprocess_user_features(u)
        assert c.dtype.is_floating_point is False and tuple(c.shape) == (1, 6) and tuple(n.shape) == (1, 13)
        cats.append(c.numpy()[0])
        nums.append(n.numpy()[0])
    np.savez_compressed(
        HERE / "preprocess_users.npz", users=json.dumps(users()),
        classes=json.dumps({c: e.classes_.tolist() for c, e in pre.label_encoders.items()}),
        scaler_mean=pre.scaler.mean_, scaler_scale=pre.scaler.scale_,
        cat=np.array(cats, dtype=np.int64), num=np.array(nums), num_dtype=str(np.array(nums).dtype))
    print("cat", np.array(cats).shape, "num", np.array(nums).dtype, np.array(nums)[0, :3])


if __name__ == "__main__":
    main()
