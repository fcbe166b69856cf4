"""rofl_crypto::bsgs32::BSGSTable (bsgs32.rs:20-73).  The table itself lives on the GPU inside the context, keyed by
its size; this object only names it."""
from . import fp


class BSGSTable:
    def __init__(self, m):                                 # BSGSTable::new, bsgs32.rs:20-34
        self.size = int(m)

    @classmethod
    def default(cls):                                      # bsgs32.rs:36-38
        return cls(1 << (fp.BSGS_N_BITS // 2 + fp.PRECOMP_BIAS))
