class _Rec:
    def __init__(self, seq):
        self.seq = seq


class _Db(dict):
    def close(self):
        pass


def index_db(idx, path, fmt):
    """Dictionary semantics (last record with an id wins); real Biopython raises on duplicates."""
    db = _Db()
    name = None
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith(">"):
                name = line[1:].split()[0]
                db[name] = _Rec("")
            elif name is not None:
                db[name].seq += line
    return db
