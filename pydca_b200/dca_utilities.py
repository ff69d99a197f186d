"""Output writers and naming of pydca/dca_utilities/dca_utilities.py that the two
command lines use.  Formats are kept byte-compatible with the reference:
score tables `'{0:<7} {1:<14} {2:<35}'` with 1-based sites (:255-266), CSV fields /
couplings (:293-360), fi / fij tables (:363-436), metadata headers (:100-172).
"""
import logging
import os

logger = logging.getLogger(__name__)

RULE = "#" + "=" * 70 + "\n"


def create_directories(dir_path):
    os.makedirs(dir_path, exist_ok=True)


def get_dca_output_file_path(output_dir, msa_file_name, prefix="", postfix=""):
    """<output_dir>/<prefix><msa base name without extension><postfix> (:29-57)."""
    root, _ = os.path.splitext(os.path.basename(msa_file_name))
    return os.path.join(output_dir, prefix.strip() + root.strip() + postfix.strip())


def mfdca_param_metadata(inst):
    return ["# PARAMETERS USED FOR THIS COMPUTATION: ",
            "#      Sequence type: {}".format(inst.biomolecule),
            "#      Total number of sequences in alignment data: {}".format(inst.num_sequences),
            "#      Length of sequences in alignment data: {}".format(inst.sequences_len),
            "#      Effective number of sequences: {}".format(inst.effective_num_sequences),
            "#      Value of sequence identity: {}".format(inst.sequence_identity),
            "#      Value of relative pseudocount: {}".format(inst.pseudocount)]


def plmdca_param_metadata(inst):
    return ["# PARAMETERS USED FOR THIS COMPUTATION: ",
            "#\tSequence type: {}".format(inst.biomolecule),
            "#\tTotal number of sequences in alignment data: {}".format(inst.num_sequences),
            "#\tLength of sequences in alignment data: {}".format(inst.sequences_len),
            "#\tValue of sequence identity: {}".format(inst.sequence_identity),
            "#\tlambda_h: {}".format(inst.lambda_h),
            "#\tlambda_J: {}".format(inst.lambda_J),
            "#\tNumber of gradient decent iterations: {}".format(inst.max_iterations)]


def mfdca_residue_repr_metadata(biomolecule):
    """Integer -> letter legend, five pairs per line (:175-204)."""
    alpha = "ACDEFGHIKLMNPQRSTVWY-" if biomolecule.strip().upper() == "PROTEIN" else "ACGU-"
    pairs = [(i + 1, c) for i, c in enumerate(alpha)]
    out = ["# RESIDUES IDENTIFICATION"]
    for k in range(len(pairs) // 5 + 1):
        out.append("".join(map(str, ["# "] + pairs[k * 5:(k + 1) * 5])))
    return out


def write_sorted_dca_scores(file_name, sorted_DI, metadata=None, score_type=None):
    logger.info("\n\tWriting DCA scores to file {}".format(file_name))
    with open(file_name, "w") as fh:
        fh.write(RULE)
        if metadata:
            for line in metadata:
                fh.write("{}\n".format(line))
        fh.write("# The First and Second columns represent sites and the"
                 "\n# Third column is {} DCA score\n".format(score_type))
        fh.write(RULE)
        for pair, score in sorted_DI:
            i, j = pair
            fh.write("{0:<7} {1:<14} {2:<35}\n".format(i + 1, j + 1, score))


def write_couplings_csv(file_name, couplings, metadata=None):
    logger.info("\n\tSaving couplings to file:\n\t{}".format(file_name))
    with open(file_name, "w") as fh:
        fh.write(RULE)
        if metadata:
            for data in metadata:
                fh.write("{}\n".format(data))
            fh.write(RULE)
        for site_pair, couplings_ij in couplings:
            fh.write("{},{}".format(site_pair[0] + 1, site_pair[1] + 1))
            for c in couplings_ij:
                fh.write(",{}".format(c))
            fh.write("\n")


def write_fields_csv(file_name, fields, metadata=None):
    logger.info("\n\tSaving fields to file:\n\t{}".format(file_name))
    with open(file_name, "w") as fh:
        fh.write("#{}\n".format(70 * "="))
        if metadata is not None:
            for data in metadata:
                fh.write("{}\n".format(data))
            fh.write("#{}\n".format(70 * "="))
            for site, site_fields in fields:
                fh.write("{}".format(site + 1))
                for fia in site_fields:
                    fh.write(",{}".format(fia))
                fh.write("\n")


def write_single_site_freqs(file_name, fi, seqs_len=None, num_site_states=None, metadata=None):
    logger.info("\n\tSaving single site frequencies to file:\n\t{}".format(file_name))
    with open(file_name, "w") as fh:
        fh.write(RULE)
        if metadata:
            for data in metadata:
                fh.write("{}\n".format(data))
            fh.write("# Below, the First integer refers to the site, the \n"
                     "# Second the residue at that site, and the Third is the \n"
                     "# frequency. Residue numbers are mapped as shown above.\n")
            fh.write(RULE)
        for i in range(seqs_len):
            for a in range(num_site_states):
                fh.write("{},{},{}\n".format(i + 1, a + 1, fi[i, a]))


def write_pair_site_freqs(file_name, fij, seqs_len=None, num_site_states=None, metadata=None):
    logger.info("\n\tSaving pair site frequencies (gaps are excluded) to file: \n\t{}".format(file_name))
    with open(file_name, "w") as fh:
        fh.write(RULE)
        if metadata:
            for data in metadata:
                fh.write("{}\n".format(data))
            fh.write("# Below, the First and Second integers refer to sites, the \n"
                     "# Third and Fourth residues, and the Last one is frequency for pairs.\n"
                     "# Residue numbers are mapped as shown above.\n")
            fh.write(RULE)
        pair_counter = 0
        for i in range(seqs_len - 1):
            for j in range(i + 1, seqs_len):
                for a in range(num_site_states - 1):
                    for b in range(num_site_states - 1):
                        fh.write("{},{},{},{},{}\n".format(i + 1, j + 1, a + 1, b + 1, fij[pair_counter, a, b]))
                pair_counter += 1
