"""duet_b200: the sv_phasing hot path of Duet (yekaizhou/duet) on B200.

Host side mirrors the reference modules (`read_file`, `sv_phasing_fn`, `sv_phasing`,
`write_file`); all compute runs in csrc/libduet_b200.so (sm_100a) through the C-ABI of
include/duet_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
