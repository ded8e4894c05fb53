"""Stand-in for the part of pdfflow madflow imports at module load (dtype helpers only).
TEST INFRASTRUCTURE ONLY -- no PDF interpolation is provided."""


def mkPDF(*a, **k):
    raise RuntimeError("tfshim: pdfflow PDFs are not available offline")
