class SeqRecord:
    def __init__(self, seq, id="", description=""):
        self.seq = seq
        self.id = id
        self.name = id
        self.description = description


class MultipleSeqAlignment(list):
    """A list of SeqRecord (what iteration over Biopython's class yields)."""

    def get_alignment_length(self):
        return len(self[0].seq) if self else 0
