"""effimvs_b200 -- B200-native (sm_100a) cost-volume hot path of Effi-MVS+.

Layout (only what the path needs):
  csrc/        hand-written CUDA kernels + the C-ABI (libeffimvs.so, include/effimvs.h)
  capi.py      ctypes binding of the C-ABI (raises if the library is missing)
  ops.py       torch.library custom ops ``effimvs::*`` over device pointers
  hotpath.py   CudaHotPath: the hot-path table used by net.py and dropin.py
  dropin.py    upstream-named call sites (homo_warping_new, pro_bilinear_sampler, ...) + patch(model)
  net.py       stock-PyTorch host network (FPN, GRU, cascade), checkpoint compatible
  fusion.py    geometric-consistency filter (misc/fusion.py call sites)
  scene.py     per-scene runner: reference views sharded over ranks, NCCL all-gather
  synthetic.py seeded synthetic inputs at DTU / T&T shapes
"""
PACKAGE_DIR = __path__[0]
