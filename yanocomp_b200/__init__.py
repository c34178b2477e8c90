'''
yanocomp_b200 -- B200 (sm_100a) implementation of the `yanocomp gmmtest`
per-position hot path (KS filter -> GMM fit -> G-test -> single-molecule
posteriors) behind the reference's own function boundary
(`parallel_test`, reference yanocomp/gmmtest.py:182).

The numerical work is done by hand-written CUDA kernels in
`csrc/libyanocomp_b200.so` (C ABI: include/yanocomp_b200.h); there is no CPU
fallback: importing `yanocomp_b200.engine` without the built library, or
running it without a GPU, raises.
'''
__version__ = '0.1.0'
