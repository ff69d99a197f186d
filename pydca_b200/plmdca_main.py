"""`plmdca` command line -- same sub-commands, flags, defaults and output files as
pydca/plmdca_main.py:259-352 (`plmdca {compute_fn,compute_di,compute_params}`).

    python -m pydca_b200.plmdca_main compute_fn protein msa.fa --max_iterations 500 --apc

--num_threads is accepted and ignored (the work runs on the GPU).  --refseq_file maps
the scores to the positions of a reference sequence through pydca_b200.sequence_backmapper
(Biopython-free; the search over the alignment runs on the GPU).
"""
import logging
import os
import sys
from argparse import ArgumentParser

from . import dca_utilities
from .plmdca import PlmDCA

logger = logging.getLogger(__name__)

DCA_COMPUTATION_SUBCOMMANDS = ("compute_fn", "compute_di", "compute_params")


def configure_logging():
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")


def _seqbackmapper(msa_file, refseq_file, biomolecule):
    """plmdca_main.py:185-192: back-mapping to a reference sequence when --refseq_file is given."""
    if not refseq_file:
        return None
    from .sequence_backmapper import SequenceBackmapper
    return SequenceBackmapper(msa_file=msa_file, refseq_file=refseq_file, biomolecule=biomolecule)


def execute_from_command_line(biomolecule, msa_file, the_command=None, refseq_file=None, seqid=None,
                              lambda_h=None, lambda_J=None, max_iterations=None, apc=False, verbose=False,
                              output_dir=None, num_threads=None, ranked_by=None, linear_dist=None,
                              num_site_pairs=None, arithmetic=None, estimator=None):
    """pydca/plmdca_main.py:136-256."""
    if verbose:
        configure_logging()
    inst = PlmDCA(msa_file, biomolecule, seqid=seqid, lambda_h=lambda_h, lambda_J=lambda_J,
                  max_iterations=max_iterations, num_threads=num_threads, verbose=verbose, arithmetic=arithmetic,
                  estimator=estimator)
    if the_command not in DCA_COMPUTATION_SUBCOMMANDS:
        return None
    mapper = _seqbackmapper(msa_file, refseq_file, biomolecule)
    meta = dca_utilities.plmdca_param_metadata(inst)
    if not output_dir:
        base, _ = os.path.splitext(os.path.basename(msa_file))
        output_dir = "PLMDCA_output_" + base
    dca_utilities.create_directories(output_dir)

    def out(prefix):
        return dca_utilities.get_dca_output_file_path(output_dir, msa_file, prefix=prefix, postfix=".txt")

    if the_command == "compute_fn":
        if apc:
            score_type = "PLMDCA Frobenius norm, average product corrected (APC)"
            scores, path = inst.compute_sorted_FN_APC(seqbackmapper=mapper), out("PLMDCA_apc_fn_scores_")
        else:
            score_type = "PLMDCA Frobenius norm, non-APC (not average product corrected)"
            scores, path = inst.compute_sorted_FN(seqbackmapper=mapper), out("PLMDCA_raw_fn_scores_")
        dca_utilities.write_sorted_dca_scores(path, scores, metadata=meta, score_type=score_type)
    if the_command == "compute_di":
        if apc:
            score_type = "PLMDCA  DI scores, average product corrected (APC)"
            scores, path = inst.compute_sorted_DI_APC(seqbackmapper=mapper), out("PLMDCA_apc_di_scores_")
        else:
            score_type = "PLMDCA DI scores, non-APC (not average product corrected)"
            scores, path = inst.compute_sorted_DI(seqbackmapper=mapper), out("PLMDCA_raw_di_scores_")
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
    return None


def build_parser():
    parser = ArgumentParser(prog="plmdca")
    sub = parser.add_subparsers(dest="subcommand_name")
    helps = {"compute_fn": "Computes DCA scores summarized by the Frobenius norm of couplings.",
             "compute_di": "Computes DCA scores summarized by direct information.",
             "compute_params": "Computes the fields and couplings of the plmDCA model."}
    for name in DCA_COMPUTATION_SUBCOMMANDS:
        p = sub.add_parser(name, help=helps[name])
        p.add_argument("biomolecule", help="Type of biomolecule (protein or RNA)")
        p.add_argument("msa_file", help="Multiple sequence alignment (MSA) file in FASTA format")
        p.add_argument("--seqid", type=float, help="Cut-off value of sequences similarity (default 0.8)")
        p.add_argument("--lambda_h", type=float, help="Fields penalizing constant (default 0.2(L-1))")
        p.add_argument("--lambda_J", type=float, help="Couplings penalizing constant (default 0.2(L-1))")
        p.add_argument("--max_iterations", type=int, help="Maximum number of L-BFGS iterations (default 100)")
        p.add_argument("--num_threads", type=int, help="Accepted for compatibility; the work runs on the GPU")
        p.add_argument("--refseq_file", help="FASTA file containing a reference sequence")
        p.add_argument("--verbose", action="store_true", help="Show logging information on the terminal")
        p.add_argument("--output_dir", help="Directory path to which output results are written")
        p.add_argument("--arithmetic", choices=("accurate", "reference"),
                       help="accurate: float64 reductions (default, fastest); reference: the reference's sequential "
                            "float32 sums bit for bit, i.e. its own L-BFGS trajectory (not a flag of the reference)")
        p.add_argument("--estimator", choices=("joint", "asymmetric"),
                       help="joint: the reference's fit of one symmetric coupling matrix (default); asymmetric: every "
                            "site fits its own couplings independently, sites spread over the GPUs without any "
                            "exchange, symmetrised at the end (a different estimator; not a flag of the reference)")
        if name != "compute_params":
            p.add_argument("--apc", action="store_true", help="Compute the average product corrected (APC) scores")
        else:
            p.add_argument("--ranked_by", choices=("FN", "FN_APC", "DI", "DI_APC", "fn", "fn_apc", "di", "di_apc"),
                           help="Method in which DCA scores are ranked")
            p.add_argument("--linear_dist", type=int, help="Minimum separation |i - j| between site pairs")
            p.add_argument("--num_site_pairs", type=int, help="Number of site pairs to extract couplings for")
    return parser


def run_plm_dca(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    args = vars(parser.parse_args(argv if argv else ["--help"]))
    execute_from_command_line(args.get("biomolecule"), args.get("msa_file"),
                              the_command=args.get("subcommand_name"), refseq_file=args.get("refseq_file"),
                              seqid=args.get("seqid"), lambda_h=args.get("lambda_h"), lambda_J=args.get("lambda_J"),
                              max_iterations=args.get("max_iterations"), num_threads=args.get("num_threads"),
                              apc=args.get("apc"), output_dir=args.get("output_dir"), verbose=args.get("verbose"),
                              ranked_by=args.get("ranked_by"), linear_dist=args.get("linear_dist"),
                              num_site_pairs=args.get("num_site_pairs"), arithmetic=args.get("arithmetic"),
                              estimator=args.get("estimator"))
    return None


if __name__ == "__main__":
    run_plm_dca()
