"""Command-line flags of the reference's scripts (trainer/options.py:3-37).

``parse_args()`` returns the *parser* (not the parsed namespace), exactly like the reference — ``scripts/demo.py:251-252``
does ``parser = parse_args(); args = parser.parse_args()`` — with every flag, type and default of the reference, so a
``demo.py``-style command line parses to the same namespace.
"""
from argparse import ArgumentParser

# (flag, kwargs) in the reference's order; defaults are the reference's (trainer/options.py:5-35)
_FLAGS = (
    ("--gpu", dict(default=0, type=int)),
    ("--save_dir", dict(default="experiments", type=str)),
    ("--exp_name", dict(default="smplx_S2G", type=str)),
    ("--speakers", dict(nargs="+")),
    ("--seed", dict(default=1, type=int)),
    ("--model_name", dict(type=str)),
    ("--use_template", dict(action="store_true")),
    ("--template_length", dict(default=0, type=int)),
    ("--resume", dict(action="store_true")),
    ("--pretrained_pth", dict(default=None, type=str)),
    ("--style_layer_norm", dict(action="store_true")),
    ("--config_file", dict(default="./config/style_gestures.json", type=str)),
    ("--audio_file", dict(default=None, type=str)),
    ("--id", dict(default=0, type=int, help="0=oliver, 1=chemistry, 2=seth, 3=conan")),
    ("--only_face", dict(action="store_true")),
    ("--stand", dict(action="store_true")),
    ("--whole_body", dict(action="store_true")),
    ("--num_sample", dict(default=1, type=int)),
    ("--face_model_name", dict(default="s2g_face", type=str)),
    ("--face_model_path", dict(default="./experiments/2022-10-15-smplx_S2G-face-3d/ckpt-99.pth", type=str)),
    ("--body_model_name", dict(default="s2g_body_pixel", type=str)),
    ("--body_model_path", dict(default="./experiments/2022-11-02-smplx_S2G-body-pixel-3d/ckpt-99.pth", type=str)),
    ("--infer", dict(action="store_true")),
)


def parse_args():
    parser = ArgumentParser()
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    return parser
