"""Input side of the path: the part of the reference's desman/Variant_Filter.py that bin/desman needs
to turn a `.freq` / `sel_var.csv` table into the count tensor snps[V,S,4] (Variant_Filter.py:70-118)
and to pick the `-r` random subset (:392-410).  The likelihood-ratio variant caller (`-f`,
:320-390) is upstream of the Gibbs hot path and is not re-implemented (out of scope).
"""
import numpy as np


class Variant_Filter():

    def __init__(self, variants, randomState, optimise=True, threshold=3.84, min_coverage=5.0, qvalue_cutoff=0.1,
                 max_iter=100, min_p=0.01, mCogFilter=2.0, cogSampleFrac=0.95, Nthreshold=10):
        m = variants.to_numpy()
        self.genes = list(variants.index)
        self.position = m[:, 0]
        m = np.delete(m, 0, 1)
        if m.shape[1] % 4:
            raise ValueError("expected 4 count columns (A,C,G,T) per sample after the Position column")
        snps = np.reshape(m, (m.shape[0], m.shape[1] // 4, 4))
        vs_mean = np.mean(snps.sum(axis=2), axis=0)
        self.randomState = randomState
        self.sample_filter = vs_mean > min_coverage                     # samples below min coverage are dropped
        self.sample_indices = np.where(self.sample_filter)[0].tolist()
        self.snps_filter = snps[:, self.sample_filter, :]
        self.V = self.snps_filter.shape[0]
        self.S = self.snps_filter.shape[1]
        self.freq = self.snps_filter.sum(axis=1)
        self.threshold = threshold
        self.qvalue_cutoff = qvalue_cutoff
        self.optimise = optimise
        self.filtered = np.zeros((self.V), dtype=bool)
        self.eta = 0.96 * np.identity((4)) + 0.01 * np.ones((4, 4))
        self.NS = self.V
        self.selected = np.ones((self.V), dtype=bool)
        self.selected_indices = np.where(self.selected)[0].tolist()
        self.randomSelect = False

    def select_Random(self, random_select):
        if random_select < self.NS:
            self.randomSelect = True
            select = np.sort(self.randomState.choice(self.NS, random_select, replace=False))
            self.snps_filter_original = np.copy(self.snps_filter)
            self.snps_filter = self.snps_filter[select, :, :]
            self.NS = random_select
            self.selected_indices_original = np.copy(self.selected_indices)
            self.selected_indices = [self.selected_indices[i] for i in select]
            self.selected_original = np.copy(self.selected)
            self.selected = np.zeros((self.V), dtype=bool)
            self.selected[self.selected_indices] = True
        return self.snps_filter

    def get_filtered_VariantsLogRatio(self):
        raise NotImplementedError("variant calling (-f, Variant_Filter.py:320-390) is outside the Gibbs hot path; "
                                  "filter with the reference's Variant_Filter.py and pass the selected variants")
