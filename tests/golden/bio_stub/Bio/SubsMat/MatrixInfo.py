blosum62 = {}
