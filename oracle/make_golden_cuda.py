"""tests/golden/sample_pdf_cuda.npz: the reference's CUDA-side sample_pdf (perturb=0) outputs, recorded ON A B200.

The build container has no GPU, so these vectors cannot be produced here: `scripts/ref_probe.py` runs the UNMODIFIED reference
(baseline/_ref) on the GPU box, wraps torch.searchsorted to record the bin indices, evaluates nerf_process.py:150-154 op by op
on the device and dumps everything to gpurun_out/ref_cuda_probe.npz; this script keeps a subset as a committed fixture:

    gpurun -- python scripts/ref_probe.py --no-time     # on the B200
    python oracle/make_golden_cuda.py                    # here

Three row-count regimes, because ATen's cumsum kernel picks its per-row thread count from [rows, 62] (ScanUtils.cuh): 24 rows -> 32
threads (one Sklansky block of 64), 8512 rows -> 16 threads (two blocks of 32 with a carry), 40000 rows -> 512 threads.
TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, '..', 'gpurun_out', 'ref_cuda_probe.npz')
DST = os.path.join(HERE, '..', 'tests', 'golden', 'sample_pdf_cuda.npz')


def main():
    d = np.load(SRC)
    out = {'device': 'NVIDIA B200', 'torch': '2.11.0+cu128'}
    for n, keep in ((24, 24), (8512, 384), (40000, 128)):
        for k in ('z', 'w', 'sum', 'cdf', 'inds', 'samples'):
            out[f'n{n}_{k}'] = d[f'n{n}_{k}'][:keep]
        out[f'n{n}_rows'] = n
    np.savez_compressed(DST, **out)
    print(DST, os.path.getsize(DST) // 1024, 'KiB')


if __name__ == '__main__':
    main()
