"""Result writers with the reference's file names and layouts (desman/Output_Results.py:28-208), so
resolvenhap.py, validateSNP2.py and GeneAssign.py read the GPU results unchanged."""
import logging
import os
import sys

import numpy as np
import pandas as p


def rchop(thestring, ending):
    return thestring[:-len(ending)] if thestring.endswith(ending) else thestring


class Output_Results():

    def __init__(self, outputDir):
        self.outputDir = outputDir
        if not os.path.exists(outputDir):
            os.makedirs(outputDir)
        self.log_file_name = self.outputDir + "/log_file.txt"
        logging.basicConfig(filename=self.log_file_name, level=logging.INFO, filemode='w',
                            format='%(asctime)s:%(levelname)s:%(name)s:%(message)s')
        logging.info("Results created in {0}".format(os.path.abspath(self.outputDir)))
        print("Up and running. Check {0} for progress".format(os.path.abspath(self.log_file_name)), file=sys.stderr)

    def set_Variants(self, variants):
        self.variants = variants
        self.contig_names = variants.index.tolist()
        self.position = variants['Position']

    def set_Variant_Filter(self, variantFilter):
        self.variantFilter = variantFilter
        self.filtered_contig_names = [self.contig_names[i] for i in variantFilter.selected_indices]
        self.filtered_position = [self.position.iloc[i] for i in variantFilter.selected_indices]

    def _fit(self, name, haplo_SNP, genomes):
        with open(self.outputDir + "/" + name, "w") as f:                       # Fit,<G asked>,<G kept>,<lp*>,<mean deviance>
            f.write("Fit,%d,%d,%f,%f\n" % (genomes, haplo_SNP.G, haplo_SNP.lp_star, haplo_SNP.meanDeviance()))

    def set_haplo_SNP(self, haplo_SNP, genomes):
        self.haplo_SNP = haplo_SNP
        self._fit("fit.txt", haplo_SNP, genomes)
        logging.info("Wrote fit stats")

    def outPredFit(self, haplo_SNP, genomes):
        self._fit("fitP.txt", haplo_SNP, genomes)
        logging.info("Wrote pred fit stats")

    def _tau_frame(self, tau, names, positions, fname):
        V = len(names)
        df = p.DataFrame(np.reshape(tau, (V, self.haplo_SNP.G * 4)), index=names)
        df['Position'] = positions
        cols = df.columns.tolist()
        df = df[cols[-1:] + cols[:-1]]                                          # Position first
        df.to_csv(self.outputDir + "/" + fname)

    def output_Filtered_Tau(self, tau):
        self._tau_frame(tau, self.filtered_contig_names, self.filtered_position, "Filtered_Tau_star.csv")
        logging.info("Wrote filtered tau star haplotype predictions")

    def output_Tau_Mean(self, tauProb):
        self._tau_frame(tauProb, self.filtered_contig_names, self.filtered_position, "Tau_Mean.csv")
        logging.info("Wrote probabilistic tau haplotype predictions")

    def output_collated_Tau(self, haplo_SNP_NS, full_variants):
        sel = np.asarray(self.variantFilter.selected, dtype=bool)
        VS = haplo_SNP_NS.V + self.haplo_SNP.V
        G = self.haplo_SNP.G
        collateTau = np.zeros((VS, G, 4), dtype=np.int64)
        collatePTau = np.zeros((VS, G, 4))
        collateTau[~sel[:VS]] = haplo_SNP_NS.tau_star
        collateTau[sel[:VS]] = self.haplo_SNP.tau_star
        collatePTau[~sel[:VS]] = haplo_SNP_NS.probabilisticTau()
        collatePTau[sel[:VS]] = self.haplo_SNP.probabilisticTau()
        names = full_variants.index.tolist()
        pos = full_variants['Position']
        onames = [names[i] for i in self.variantFilter.selected_indices_original]
        opos = [pos.iloc[i] for i in self.variantFilter.selected_indices_original]
        self._tau_frame(collateTau, onames, opos, "Collated_Tau_star.csv")
        logging.info("Wrote all tau haplotype predictions")
        self._tau_frame(collatePTau, onames, opos, "Collated_Tau_mean.csv")
        logging.info("Wrote all probabilistic tau haplotype predictions")

    def _sample_names(self):
        cols = self.variants.columns.values.tolist()
        originalS = (len(cols) - 1) // 4
        names = [rchop(cols[i], '-A') for i in range(1, originalS * 4, 4)]
        return [names[i] for i in self.variantFilter.sample_indices]

    def output_Gamma_Mean(self, gamma):
        p.DataFrame(gamma, index=self._sample_names()).to_csv(self.outputDir + "/Gamma_mean.csv")
        logging.info("Wrote mean gamma haplotype relative frequencies")

    def output_Gamma(self, gamma):
        p.DataFrame(gamma, index=self._sample_names()).to_csv(self.outputDir + "/Gamma_star.csv")
        logging.info("Wrote gamma haplotype relative frequencies")

    def output_Eta(self, eta):
        p.DataFrame(eta).to_csv(self.outputDir + "/Eta_star.csv")
        logging.info("Wrote transition error matrix")

    def output_Eta_Mean(self, eta):
        p.DataFrame(eta).to_csv(self.outputDir + "/Eta_mean.csv")
        logging.info("Wrote transition error matrix")

    def output_Selected_Variants(self):
        self.variants[self.variantFilter.selected].to_csv(self.outputDir + "/Selected_variants.csv")
        logging.info("Wrote selected variants")
