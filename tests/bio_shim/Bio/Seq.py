from oracle import oracle as _O


def translate(sequence, table=1, **kw):
    return _O.translate(str(sequence), int(table))


def reverse_complement(sequence):
    return _O.revcomp(str(sequence))
