from .Align import MultipleSeqAlignment, SeqRecord


def read(handle, fmt):
    """Bio.AlignIO.read(file, 'fasta'): one alignment, records in file order, wrapped lines joined;
    ValueError for an empty file or ragged lengths (as Biopython)."""
    if fmt != "fasta":
        raise ValueError("stub supports FASTA only")
    close = False
    if isinstance(handle, str):
        handle = open(handle)
        close = True
    records, name, chunks = [], None, []
    try:
        for line in handle:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith(">"):
                if name is not None:
                    records.append(SeqRecord("".join(chunks), name.split()[0] if name.split() else "", name))
                name, chunks = line[1:], []
            elif name is not None:
                chunks.append(line.strip())
        if name is not None:
            records.append(SeqRecord("".join(chunks), name.split()[0] if name.split() else "", name))
    finally:
        if close:
            handle.close()
    if not records:
        raise ValueError("No records found in handle")
    if len({len(r.seq) for r in records}) != 1:
        raise ValueError("Sequences must all be the same length")
    return MultipleSeqAlignment(records)
