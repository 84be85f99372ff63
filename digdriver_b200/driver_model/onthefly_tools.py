"""Drop-in for DIGDriver/driver_model/onthefly_tools.py: the whole hot path fused for an ad-hoc BED file
or region string (DigDriver.py quickDriver) -- K4 (block contexts), K2 (window contexts), K6 (transfer),
K5 (observed counts) and K7 (burden test) back to back on the GPU."""
import os
import tempfile

import pandas as pd

from .. import storage
from ..data_tools import mutation_tools
from ..sequence_model import genic_driver_tools, sequence_tools
from . import transfer_tools


def region_str_to_params(region_str):
    col_split = region_str.split(":")
    chrom = col_split[0][3:] if col_split[0].startswith("chr") else col_split[0]
    pos_split = col_split[1].split("-")
    return chrom, int(pos_split[0]), int(pos_split[1])


def DIG_onthefly(f_pretrained, f_mut, f_fasta, f_elts_bed=None, region_str=None, scale_factor=None,
                 scale_factor_indel=None, scale_type="genome", scale_by_expectation=True, max_muts_per_sample=3e9,
                 max_muts_per_elt_per_sample=3e9, skip_pvals=False):
    """Reference onthefly_tools.py:28-190."""
    assert f_elts_bed or region_str, "ERROR: you must provide --f-bed or --region_str."
    temp_name = None
    if region_str:
        temp_file, temp_name = tempfile.mkstemp()
        CHROM, START, END = region_str_to_params(region_str)
        os.write(temp_file, "{}\t{}\t{}\tUserELT\t0\t+\t0\t0\t.\t1\t{},\t0,".format(CHROM, START, END, END - START).encode())
        os.close(temp_file)
        f_elts_bed = temp_name
    try:
        print('Tabulating mutations')
        df_mut_tab, blacklist = mutation_tools.tabulate_mutations_in_element(
            f_mut, f_elts_bed, bed12=True, drop_duplicates=True, all_elements=True,
            max_muts_per_sample=max_muts_per_sample, max_muts_per_elt_per_sample=max_muts_per_elt_per_sample,
            return_blacklist=True)
        if scale_by_expectation:
            cj, cj_indel = transfer_tools._expectation_scale_factors(f_mut, f_pretrained, blacklist)
        elif scale_factor:
            cj, cj_indel = scale_factor, scale_factor_indel
        else:
            print('Calculating scale factor')
            cj, cj_indel = transfer_tools.calc_scale_factor_efficient(f_mut, f_pretrained, scale_type=scale_type)

        # K4: strand-aware context counts of every block (the reference hard-codes 10 processes / 10 kb here)
        L_contexts = sequence_tools.precount_region_contexts_parallel(f_elts_bed, f_fasta, 10, 10000, sub_elts=True,
                                                                      n_up=1, n_down=1)
        pre = storage.Store(f_pretrained, "r")
        rm = genic_driver_tools.RegionModel(pre.read_table('region_params'))
        d_pr = sequence_tools.d_pr_from_model192(pre.read_table('sequence_model_192'))
        df_elts = mutation_tools.bed12_boundaries(f_elts_bed)
        # K2 on the windows of the region model (the reference counts them per element from the FASTA), K6
        win_counts = genic_driver_tools._window_counts_for(rm, f_fasta)
        df_pre = genic_driver_tools.nonc_model_arrays(df_elts, L_contexts, rm, win_counts, d_pr)
        alpha = df_pre.MU ** 2 / df_pre.SIGMA ** 2
        theta = df_pre.SIGMA ** 2 / df_pre.MU
        pretrain_df = pd.DataFrame({
            'ELT_SIZE': df_pre.ELT_SIZE.values, 'FLAG': df_pre.FLAG.values, 'R_SIZE': df_pre.R_SIZE.values,
            'R_OBS': df_pre.R_OBS.values, 'R_INDEL': df_pre.R_INDEL.values, 'MU': df_pre.MU.values,
            'SIGMA': df_pre.SIGMA.values, 'ALPHA': alpha.values, 'THETA': (theta * cj).values,
            'MU_INDEL': df_pre.MU.values, 'SIGMA_INDEL': df_pre.SIGMA.values, 'ALPHA_INDEL': alpha.values,
            'THETA_INDEL': (theta * cj_indel).values, 'Pi_SUM': df_pre.P_SUM.values, 'Pi_INDEL': df_pre.P_INDEL.values,
        }, index=df_pre.ELT.values)
        df_model = df_mut_tab.merge(pretrain_df, left_on='ELT', right_index=True)
        df_model = transfer_tools.element_expected_muts_nb(df_model)
        if not skip_pvals:
            df_model = transfer_tools.element_pvalue_burden_nb(df_model)
            df_model = transfer_tools.element_pvalue_burden_nb_by_sample(df_model)
            # quirk kept: THETA_INDEL already carries cj_indel and element_pvalue_indel multiplies again (:151,:181)
            df_model = transfer_tools.element_pvalue_indel(df_model, cj_indel)
            df_model['PVAL_MUT_BURDEN'] = transfer_tools.fisher_combine(df_model.PVAL_SNV_BURDEN.values,
                                                                        df_model.PVAL_INDEL_BURDEN.values)
        return df_model
    finally:
        if temp_name:
            os.remove(temp_name)
