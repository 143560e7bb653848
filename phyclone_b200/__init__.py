"""phyclone_b200 -- B200-native implementation of PhyClone's tree-node likelihood hot path.

The compute lives in `csrc/` (hand-written sm_100a CUDA behind the C ABI declared in
`include/phyclone_b200.h`); this package is the Python host side that mirrors the
reference's callables for that path.  There is no CPU fallback:
every compute entry point raises if the shared library or a CUDA device is missing.
"""

__version__ = "0.1.0"
