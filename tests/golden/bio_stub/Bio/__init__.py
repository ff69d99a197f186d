"""Minimal stand-in for Biopython -- TEST INFRASTRUCTURE ONLY.

Biopython is not installed in this image and there is no network.  The reference's DCA classes
use a tiny part of it (AlignIO.read(path, 'fasta') yielding records with a .seq, and the
Align.MultipleSeqAlignment type for an isinstance test); this package provides exactly that so
that tests/golden/make_class_goldens.py can run the UNMODIFIED reference classes and command
lines.  Nothing under pydca_b200/ imports it.  pairwise2 is provided by
tests/golden/bio_stub/Bio/pairwise2.py only as far as sequence_backmapper.py needs it.
"""
