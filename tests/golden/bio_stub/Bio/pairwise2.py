"""Placeholder: importable so that pydca.sequence_backmapper can be imported; the alignment
functions are not provided (the class goldens do not use a reference sequence)."""


class _Unavailable:
    def __getattr__(self, name):
        raise NotImplementedError("Bio.pairwise2.%s is not part of the test stub" % name)


align = _Unavailable()
