"""Drop-in for the reference's config.py:18-111 option set (same names, defaults and the paired
*_true/*_false store flags), restated over argparse because configargparse is not a dependency
here.  ``--config file`` accepts the reference's ``key = value`` files (configs/blender/lego.txt)."""
import argparse
import os

LOG_DIR = os.path.join(os.path.abspath(os.path.dirname(os.path.realpath(__file__))), "logs")


def _config_file_args(path):
    out = []
    with open(path) as f:
        for line in f:
            line = line.split('#', 1)[0].strip()
            if not line:
                continue
            if '=' in line:
                k, v = [s.strip() for s in line.split('=', 1)]
                v = v.strip('[]')
                out.append('--' + k)
                out.extend(x.strip() for x in v.split(',') if x.strip()) if k == 'gpu_ids' else out.append(v)
            else:
                out.append('--' + line)
    return out


def build_parser():
    p = argparse.ArgumentParser(add_help=False)
    p.add_argument('--config', type=str, default=None, help='config file path')
    p.set_defaults(visdom=True)
    p.add_argument('--visdom_port', type=int, default=8900)
    p.add_argument('--gpu_ids', nargs='+', default=['0'])
    p.add_argument('--data_type', type=str, help='[ blender, llff, custom ]')
    p.add_argument('--data_name', type=str)
    p.add_argument('--data_root', type=str)
    p.add_argument('--downsample', type=int, default=0)
    p.add_argument('--near', type=float)
    p.add_argument('--far', type=float)
    p.set_defaults(bkg_white=False)
    p.add_argument('--bkg_white_true', dest='bkg_white', action='store_true')
    p.set_defaults(colmap_relaunch=False)
    p.add_argument('--colmap_relaunch_true', dest='colmap_relaunch', action='store_true')
    p.add_argument('--precrop_iters', type=int, default=0)
    p.add_argument('--precrop_frac', type=float, default=.5)
    p.add_argument('--video_batch', type=int)
    p.add_argument('--L_x', type=int, default=10)
    p.add_argument('--L_d', type=int, default=4)
    p.add_argument('--netDepth', type=int, default=8)
    p.add_argument('--netWidth', type=int, default=256)
    p.add_argument('--exp_name', type=str)
    p.add_argument('--lr', type=float, default=5e-4)
    p.add_argument('--lr_min', type=float, default=5e-5)
    p.add_argument('--iter_warmup', type=int, default=10000)
    p.add_argument('--iter_N', type=int)
    p.add_argument('--iter_start', type=int, default=0)
    p.set_defaults(global_batch=True)
    p.add_argument('--global_batch_false', dest='global_batch', action='store_false')
    p.add_argument('--N_rays', type=int, default=4096)
    p.add_argument('--N_samples_c', type=int, default=64)
    p.add_argument('--N_samples_f', type=int, default=128)
    p.add_argument('--chunk_rays', type=int, default=4096)
    p.add_argument('--chunk_pts', type=int, default=524288)
    p.add_argument('--perturb', default=1.)       # untyped in the reference too (config.py:76, SURVEY B-4)
    p.set_defaults(mode_test=True)
    p.add_argument('--mode_test_false', dest='mode_test', action='store_false')
    p.add_argument('--testskip', type=int)
    p.set_defaults(mode_render=True)
    p.add_argument('--mode_render_false', dest='mode_render', action='store_false')
    p.add_argument('--render_type', type=str, default='gif')
    p.add_argument('--n_angle', type=int)
    p.add_argument('--single_angle', type=float, default=-1)
    p.add_argument('--phi', type=float)
    p.add_argument('--nf', type=float)
    p.add_argument('--testing_idx', type=int)
    p.add_argument('--idx_vis', type=int, default=100)
    p.add_argument('--idx_print', type=int, default=1000)
    p.add_argument('--idx_save', type=int)
    p.add_argument('--idx_test', type=int)
    p.add_argument('--idx_render', type=int)
    p.add_argument('--idx_vis_cam_param', type=int, default=1000)
    # engine options (not in the reference)
    p.add_argument('--precision', type=str, default='bf16', choices=['fp32', 'bf16'])
    p.add_argument('--seed', type=int, default=0)
    return p


def get_args_parser(argv=None):
    """config.py:18-111 -> opts (+ world_size, rank, int gpu_ids as config.py:106-109 / main.py:20)."""
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = build_parser()
    pre, _ = parser.parse_known_args(argv)
    if pre.config:
        argv = _config_file_args(pre.config) + argv       # CLI overrides the file
    opts = parser.parse_args(argv)
    opts.gpu_ids = [int(g) for g in opts.gpu_ids]
    opts.world_size = len(opts.gpu_ids)
    opts.rank = 0
    return opts
