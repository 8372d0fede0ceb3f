"""Install the UNMODIFIED reference (the files of its ray-batch hot path) into the git-ignored baseline/_ref/.

    python baseline/install_ref.py            # build container only: needs /root/reference

The reference is a flat script tree without setup.py / pyproject.toml, so `pip install /root/reference` has nothing
to build; BASELINE.md section 4 step 1 names the files to take.  They are copied byte for byte (checked below) and
never enter git history (.gitignore: baseline/_ref/); the directory travels to the GPU box with the gpurun snapshot,
where bench.py --impl reference, the `gpu_baseline` leg and scripts/ref_probe.py import it through baseline/ref_shim.py.
"""
import filecmp
import os
import shutil
import sys

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['rays.py', 'nerf_process.py', 'utils.py', 'scheduler.py', 'model/__init__.py', 'model/NeRF.py', 'model/NeRFHelper.py',
         'model/PositionalEncoding.py', 'dataset/render_pose.py']


def install(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f'{REF} not present: keeping the existing {DST}' if os.path.isdir(DST) else f'{REF} not present and no {DST}')
        return os.path.isdir(DST)
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
        assert filecmp.cmp(os.path.join(REF, f), dst, shallow=False)
    if verbose:
        print(f'installed {len(FILES)} reference files into {DST}')
    return True


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
