"""Where the public call's time goes: FlashDeconv.fit_transform(scipy CSR in pageable memory) at C3, stage by stage."""
import time, sys, numpy as np, torch
sys.path.insert(0, ".")
from scipy import sparse
from flashdeconv_b200 import pipeline, genes, synth
from flashdeconv_b200.synth import make_dataset_device, CONFIGS

c = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
data = make_dataset_device(c["n_spots"], c["n_genes"], c["n_types"], c["depth"], jitter=c["jitter"], seed=0, device="cuda:0", pinned=False)
Y = sparse.csr_matrix((data["data"].cpu().numpy(), data["indices"].cpu().numpy(), data["indptr"].cpu().numpy()),
                      shape=(c["n_spots"], c["n_genes"]))
X, coords = data["X"], data["coords"].cpu().numpy()
del data
torch.cuda.empty_cache()

def T(f, name, rep=3):
    best = 1e9
    for _ in range(rep):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{name:34s} {1e3 * best:8.1f} ms", flush=True)
    return out

T(lambda: Y.has_canonical_format, "has_canonical_format (cached)")
csr = T(lambda: pipeline.csr_to_device(Y), "csr_to_device (pageable H2D)")
gi, lev = T(lambda: genes.select_informative_genes_device(csr, np.asarray(X), n_hvg=2000, n_markers_per_type=50), "gene selection")
res = T(lambda: pipeline.deconvolve_path(csr, X, coords, gi, lev), "deconvolve_path (device CSR)")
res = T(lambda: pipeline.deconvolve_path(csr, X, coords, gi, lev, pinned_out=True), "  same, pinned outputs")
from flashdeconv_b200 import FlashDeconv
T(lambda: FlashDeconv(random_state=0).fit_transform(Y, X, coords), "FlashDeconv.fit_transform")
