"""TEST-ONLY stand-in for Biopython (not installed in this image), just enough for the reference's
pure-Python stages (lib/aln.py prepare_target/raw2target, lib/assemble.py) to import and run when
tests/golden/make_golden.py generates golden vectors. translate()/reverse_complement() delegate to
the oracle's restatement of Biopython's semantics. Never imported by the product."""


class BiopythonWarning(Warning):
    pass
