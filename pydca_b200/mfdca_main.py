"""`mfdca` command line -- same sub-commands, flags, defaults and output files as
pydca/mfdca_main.py:299-420 (`mfdca {compute_di,compute_fn,compute_params,compute_fi,compute_fij}`).

    python -m pydca_b200.mfdca_main compute_fn rna MSA_RF00167.fa --pseudocount 0.5 --apc
"""
import logging
import os
import sys
from argparse import ArgumentParser

from . import dca_utilities
from .meanfield_dca import MeanFieldDCA
from .plmdca_main import _seqbackmapper, configure_logging

logger = logging.getLogger(__name__)

DCA_COMPUTATION_SUBCOMMANDS = ("compute_di", "compute_fn", "compute_params", "compute_fi", "compute_fij")


def execute_from_command_line(msa_file=None, biomolecule=None, seqid=None, pseudocount=None, the_command=None,
                              refseq_file=None, verbose=False, output_dir=None, apc=False, ranked_by=None,
                              linear_dist=None, num_site_pairs=None):
    """pydca/mfdca_main.py:130-296."""
    if verbose:
        configure_logging()
    if the_command is None or the_command.strip() not in DCA_COMPUTATION_SUBCOMMANDS:
        return None
    the_command = the_command.strip()
    mapper = _seqbackmapper(msa_file, refseq_file, biomolecule)
    inst = MeanFieldDCA(msa_file, biomolecule, seqid=seqid, pseudocount=pseudocount)
    meta = dca_utilities.mfdca_param_metadata(inst)
    if not output_dir:
        base, _ = os.path.splitext(os.path.basename(msa_file))
        output_dir = "MFDCA_output_" + base
    dca_utilities.create_directories(output_dir)

    def out(prefix):
        return dca_utilities.get_dca_output_file_path(output_dir, msa_file, prefix=prefix, postfix=".txt")

    if the_command == "compute_di":
        if apc:
            score_type = " MF DI average product corrected (APC)"
            scores, path = inst.compute_sorted_DI_APC(seqbackmapper=mapper), out("MFDCA_apc_di_scores_")
        else:
            score_type = "raw DI"
            scores, path = inst.compute_sorted_DI(seqbackmapper=mapper), out("MFDCA_raw_di_scores_")
        dca_utilities.write_sorted_dca_scores(path, scores, metadata=meta, score_type=score_type)
    if the_command == "compute_fn":
        if apc:
            score_type = "MFDCA Frobenius norm, average product corrected (APC)"
            scores, path = inst.compute_sorted_FN_APC(seqbackmapper=mapper), out("MFDCA_apc_fn_scores_")
        else:
            score_type = "MFDCA raw Frobenius norm"
            scores, path = inst.compute_sorted_FN(seqbackmapper=mapper), out("MFDCA_raw_fn_scores_")
        dca_utilities.write_sorted_dca_scores(path, scores, metadata=meta, score_type=score_type)
    if the_command == "compute_params":
        fields, couplings = inst.compute_params(seqbackmapper=mapper, ranked_by=ranked_by, linear_dist=linear_dist,
                                                num_site_pairs=num_site_pairs)
        meta.append("#\tTotal number of sites whose fields are extracted: {}".format(len(fields)))
        dca_utilities.write_fields_csv(out("fields_"), fields, metadata=meta)
        meta.pop()
        meta.append("#\tTotal number of site pairs whose couplings are extracted: {}".format(len(couplings)))
        meta.append("#\tDCA ranking method used: {}".format("FN_APC" if ranked_by is None else ranked_by))
        meta.append("#\tMinimum separation beteween site pairs in sequence: |i - j| > {}".format(
            4 if linear_dist is None else linear_dist))
        dca_utilities.write_couplings_csv(out("couplings_"), couplings, metadata=meta)
    if the_command == "compute_fi":         # --pseudocount 0.0 gives raw frequencies
        fi = inst.get_reg_single_site_freqs()
        metadata = meta + dca_utilities.mfdca_residue_repr_metadata(inst.biomolecule)
        dca_utilities.write_single_site_freqs(out("fi_"), fi, seqs_len=inst.sequences_len,
                                              num_site_states=inst.num_site_states, metadata=metadata)
    if the_command == "compute_fij":
        metadata = meta + dca_utilities.mfdca_residue_repr_metadata(inst.biomolecule)
        fij = inst.get_reg_pair_site_freqs()
        dca_utilities.write_pair_site_freqs(out("fij_"), fij, seqs_len=inst.sequences_len,
                                            num_site_states=inst.num_site_states, metadata=metadata)
    return None


def build_parser():
    parser = ArgumentParser(prog="mfdca")
    sub = parser.add_subparsers(dest="subcommand_name")
    helps = {"compute_di": "Computes the direct information.",
             "compute_fn": "Compute DCA score using the Frobenius norm of the couplings.",
             "compute_params": "Computes the fields and couplings of the mean-field model.",
             "compute_fi": "Computes regularized single-site frequencies from MSA (--pseudocount 0 for raw).",
             "compute_fij": "Computes regularized pair-site frequencies from MSA (--pseudocount 0 for raw)."}
    for name in DCA_COMPUTATION_SUBCOMMANDS:
        p = sub.add_parser(name, help=helps[name])
        p.add_argument("biomolecule", help="Type of biomolecule (protein or RNA)")
        p.add_argument("msa_file", help="Multiple sequence alignment (MSA) file in FASTA format")
        p.add_argument("--seqid", type=float, help="Cut-off value of sequences similarity (default 0.8)")
        p.add_argument("--pseudocount", type=float, help="Relative value of the pseudocount (default 0.5)")
        p.add_argument("--refseq_file", help="FASTA file containing a reference sequence")
        p.add_argument("--output_dir", help="Directory path to which output results are written")
        p.add_argument("--verbose", action="store_true", help="Show logging information on the terminal")
        if name != "compute_params":
            p.add_argument("--apc", action="store_true", help="Compute the average product corrected (APC) scores")
        else:
            p.add_argument("--ranked_by", choices=("FN", "FN_APC", "DI", "DI_APC", "fn", "fn_apc", "di", "di_apc"),
                           help="Method in which DCA scores are ranked")
            p.add_argument("--linear_dist", type=int, help="Minimum separation |i - j| between site pairs")
            p.add_argument("--num_site_pairs", type=int, help="Number of site pairs to extract couplings for")
    return parser


def run_meanfield_dca(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    args = vars(parser.parse_args(argv if argv else ["--help"]))
    execute_from_command_line(biomolecule=args.get("biomolecule"), msa_file=args.get("msa_file"),
                              seqid=args.get("seqid"), pseudocount=args.get("pseudocount"),
                              refseq_file=args.get("refseq_file"), the_command=args.get("subcommand_name"),
                              verbose=args.get("verbose"), output_dir=args.get("output_dir"), apc=args.get("apc"),
                              ranked_by=args.get("ranked_by"), linear_dist=args.get("linear_dist"),
                              num_site_pairs=args.get("num_site_pairs"))
    logger.info("\n\tDONE")
    return None


if __name__ == "__main__":
    run_meanfield_dca()
