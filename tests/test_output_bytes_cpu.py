"""the result writers are byte-compatible with the reference's.

The arrays the UNMODIFIED reference wrote for `desman COG0015 -g 5 -i 50` (recorded in cog0015_i50.npz by
tests/golden/make_golden.py) are fed through desman_b200.Output_Results; every file must equal, byte for byte, the file the
reference's own Output_Results (Output_Results.py:63-208) wrote in that run (tests/golden/cog0015_i50/).  No GPU involved.
"""
import os

import numpy as np
import pandas as p
from numpy.random import RandomState

from conftest import GOLDEN, golden, onehot


def _write_freq(path):
    z = golden("cog0015.npz")
    snps = z["snps"].astype(np.int64)
    cols = [str(c) for c in z["columns"]]
    with open(path, "w") as f:
        f.write("Contig," + ",".join(cols) + "\n")
        flat = snps.reshape(snps.shape[0], -1)
        for i in range(snps.shape[0]):
            f.write("%s,%d,%s\n" % (z["contigs"][i], z["position"][i], ",".join(str(x) for x in flat[i])))


class _Fitted:
    """What Output_Results reads of a fitted sampler (Output_Results.py:63-70)."""

    def __init__(self, z):
        self.G = int(z["meta"][5])
        self.V = int(z["meta"][0])
        self.lp_star = float(z["lp_star"])
        self._dev = float(z["mean_dev"])

    def meanDeviance(self):
        return self._dev


def test_writers_byte_identical_to_reference_run(tmp_path):
    from desman_b200 import Output_Results as outr
    from desman_b200 import Variant_Filter as vf
    z = golden("cog0015_i50.npz")
    freq = tmp_path / "cog0015.freq"
    _write_freq(str(freq))
    variants = p.read_csv(str(freq), header=0, index_col=0)                  # bin/desman:84
    filt = vf.Variant_Filter(variants, randomState=RandomState(238329), optimise=True, threshold=None, min_coverage=5.0,
                             qvalue_cutoff=1.0e-3)
    assert np.array_equal(filt.snps_filter, golden("cog0015.npz")["snps"])    # the loader's tensor is the recorded one
    out = tmp_path / "out"
    w = outr.Output_Results(str(out))
    w.set_Variants(variants)
    w.set_Variant_Filter(filt)
    w.set_haplo_SNP(_Fitted(z), 5)
    w.output_Filtered_Tau(onehot(z["tau_star"]))
    w.output_Tau_Mean(z["tau_mean"])
    w.output_Gamma(z["gamma_star"])
    w.output_Gamma_Mean(z["gamma_mean"])
    w.output_Eta(z["eta_star"])
    w.output_Eta_Mean(z["eta_mean"])
    w.output_Selected_Variants()
    ref_dir = os.path.join(GOLDEN, "cog0015_i50")
    for name in ("fit.txt", "Filtered_Tau_star.csv", "Tau_Mean.csv", "Gamma_star.csv", "Gamma_mean.csv", "Eta_star.csv",
                 "Eta_mean.csv"):
        mine = open(out / name, "rb").read()
        ref = open(os.path.join(ref_dir, name), "rb").read()
        assert mine == ref, name
    # every position is selected without -f / -r: Selected_variants.csv is the input table written back (Output_Results.py:205-208)
    assert open(out / "Selected_variants.csv").read() == open(freq).read()
