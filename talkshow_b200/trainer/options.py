"""Command-line flags of the reference's scripts (trainer/options.py:3-37), inference subset kept
verbatim in name/type/default so ``scripts/demo.py``-style invocations parse unchanged."""
import argparse


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpu", default=0, type=int)
    p.add_argument("--save_dir", default="experiments", type=str)
    p.add_argument("--exp_name", default="smplx_S2G", type=str)
    p.add_argument("--speakers", nargs="+")
    p.add_argument("--seed", default=1, type=int)
    p.add_argument("--model_name", type=str)
    p.add_argument("--config_file", default="./config/body_pixel.json", type=str)
    p.add_argument("--face_model_name", type=str)
    p.add_argument("--face_model_path", type=str)
    p.add_argument("--body_model_name", type=str)
    p.add_argument("--body_model_path", type=str)
    p.add_argument("--audio_file", default=None, type=str)
    p.add_argument("--id", default=0, type=int, help="0=oliver, 1=chemistry, 2=seth, 3=conan")
    p.add_argument("--only_face", action="store_true")
    p.add_argument("--stand", action="store_true")
    p.add_argument("--whole_body", action="store_true")
    p.add_argument("--num_sample", default=1, type=int)
    p.add_argument("--infer", action="store_true")
    return p.parse_args(argv)
