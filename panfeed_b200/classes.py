"""Host-side records, same fields as the reference's namedtuples
(`/root/reference/panfeed/classes.py:5-18`) so callers can be switched over."""
from collections import namedtuple

Feature = namedtuple("Feature", ["id", "chromosome", "start", "end", "strand"])

Seqinfo = namedtuple("Seqinfo", ["sequence", "compsequence", "id", "chromosome",
                                 "start", "end", "strand", "offset"])
