"""Golden vectors for SURVEY.md 8f rank 4: runs the UNMODIFIED ``compute_vectorfield_features`` of the reference
(/root/reference/examples/eeg_example/eeg_utils.py:46-80) on a seeded input.  The module itself cannot be imported here
(it imports mne / matplotlib at top level), so the function's source is extracted with ``ast`` and executed as is with the
two names it needs (np, KDTree).  Run in the build container only:  python tests/golden/make_golden_eeg.py"""
import ast
import os

import numpy as np
from scipy.spatial import KDTree

REF = "/root/reference/examples/eeg_example/eeg_utils.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_function():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_vectorfield_features"][0]
    ns = {"np": np, "KDTree": KDTree}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["compute_vectorfield_features"]


def main():
    f = reference_function()
    rng = np.random.default_rng(0)
    X = rng.normal(size=(300, 3))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X = X[X[:, 2] > -0.2]                                   # scalp-like cap
    V = rng.normal(size=X.shape)
    out = {"positions": X, "vectors": V}
    for k in (5, 3):
        div, curl = f(X, V, k=k)
        out["div_k%d" % k], out["curl_k%d" % k] = div, curl
    np.savez_compressed(os.path.join(HERE, "eeg_features.npz"), **out)
    print("wrote eeg_features.npz", X.shape)


if __name__ == "__main__":
    main()
