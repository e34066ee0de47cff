"""scripts/demo.py of the reference on the CUDA engine: ``python -m talkshow_b200.scripts.demo --config_file
config/body_pixel.json --infer --audio_file x.wav --body_model_path ... --face_model_path ... --num_sample N``.

Same flow as the reference's ``infer`` (scripts/demo.py:158-246): one deterministic face pass, ``num_sample``
body samples, jaw | body+hands | expression -> ``part2full`` -> ``(num_sample * F, 265)`` saved as
``visualise/video/<config.Log.name>/<wav stem>.npy``.  Differences, all on purpose:
  * the ``num_sample`` body chains run as ONE batched engine call.  The sampler noise is still drawn sample by sample
    in the reference's order ([1,2048] per draw, all 2T draws of sample 0 first), so under the same seed and generator
    device the sampled codes equal the reference's sequential loop;
  * SMPL-X vertices come from ``talkshow_b200.smplx_lbs`` (batched LBS on the device) when a model dict is given;
    mesh rendering (pyrender / ffmpeg, visualise/rendering.py) is out of scope and skipped.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..data_utils.lower_body import part2full
from ..nets import s2g_body_pixel, s2g_body_vq, s2g_face
from ..nets.base import draw_sampler_noise
from ..trainer.config import load_JsonConfig
from ..trainer.options import parse_args


def init_model(model_name, model_path, args, config):
    """scripts/demo.py:30-62: wrapper by name + checkpoint in any of the three layouts the reference accepts."""
    table = {"s2g_face": s2g_face, "s2g_body_vq": s2g_body_vq, "s2g_body_pixel": s2g_body_pixel}
    if model_name not in table:
        raise NotImplementedError(model_name)
    generator = table[model_name](args, config)
    ckpt = torch.load(model_path, map_location=torch.device("cpu"))
    if "generator" in ckpt:
        generator.load_state_dict(ckpt["generator"])
    else:
        generator.load_state_dict({"generator": ckpt})
    return generator


def npy_path(config, wav_file):
    """scripts/demo.py:241-243."""
    stem = wav_file.split("\\")[-1].split(".")[-2].split("/")[-1]
    return "visualise/video/" + config.Log.name + "/" + stem


def infer(g_body, g_face, smplx_model, rendertool, config, args, save=True):
    """-> (result_list: num_sample tensors [F,265] on the device, vertices_list or None)."""
    device = g_body.device
    num_sample, wav, stand = args.num_sample, args.audio_file, args.stand
    pred_face = torch.as_tensor(g_face.infer_on_audio(wav, initial_pose=None, norm_stats=None, w_pre=False, frame=None,
                                                      am=True, am_sr=16000)).squeeze(0).to(device)       # [F,103]
    if config.Data.pose.convert_to_6d:
        jaw6 = pred_face[:, :6].reshape(pred_face.shape[0], -1, 6)
        pred_jaw = g_body.engine.rot6d_to_axis_angle(jaw6).reshape(pred_face.shape[0], -1)
        pred_face = pred_face[:, 6:]
    else:
        pred_jaw, pred_face = pred_face[:, :3], pred_face[:, 3:]
    F = pred_face.shape[0]
    ident = torch.tensor([args.id])
    # one batched call for all diversity samples; noise stream = the reference's per-sample loop (see module docstring)
    noise_fn = lambda T, B: torch.cat([draw_sampler_noise(T, 1, g_body.noise_device, per_step=g_body.noise_per_step)
                                       for _ in range(B)], 1)
    pred_all = torch.as_tensor(g_body.infer_on_audio(wav, initial_pose=None, norm_stats=None, txgfile=None, id=ident,
                                                     var=None, fps=30, w_pre=False, B=num_sample, noise_fn=noise_fn)).to(device)
    result_list = []
    for i in range(num_sample):
        pred = pred_all[i]
        if pred.shape[0] < F:                                           # :207-211
            pred = torch.cat([pred, pred[-1:].repeat(F - pred.shape[0], 1)], 0)
        else:
            pred = pred[:F]
        if config.Data.pose.convert_to_6d:
            pred = g_body.engine.rot6d_to_axis_angle(pred.reshape(pred.shape[0], -1, 6)).reshape(pred.shape[0], -1)
        pred = part2full(torch.cat([pred_jaw, pred, pred_face], -1), stand)
        if args.only_face:                                              # :226-227
            static = torch.zeros(1, 162, device=device)
            static[:, 6:9] = torch.tensor([3.0747, -0.0158, -0.0152], device=device)
            pred = torch.cat([pred[:, :3], static.repeat(pred.shape[0], 1), pred[:, -100:]], -1)
        result_list.append(pred)
    vertices_list = None
    if smplx_model is not None:
        from ..smplx_lbs import get_vertices

        vertices_list = get_vertices(smplx_model, None, result_list, config.Data.pose.expression, engine=g_body.engine)
    if save:
        out = np.concatenate([r.cpu().numpy() for r in result_list], axis=0)      # (num_sample*F, 265), :239-245
        fn = npy_path(config, wav)
        os.makedirs(os.path.dirname(fn), exist_ok=True)
        np.save(fn, out)
    if rendertool is not None:
        rendertool._render_sequences(wav, vertices_list, stand=stand, face=args.only_face, whole_body=args.whole_body)
    return result_list, vertices_list


def main(argv=None):
    parser = parse_args()
    args = parser.parse_args(argv)
    config = load_JsonConfig(args.config_file)
    print("init model...")
    g_body = init_model(args.body_model_name, args.body_model_path, args, config)
    g_face = init_model(args.face_model_name, args.face_model_path, args, config)
    smplx_model = None
    npz = getattr(config, "smplx_npz_path", None)
    if npz and os.path.exists(npz):
        from ..smplx_lbs import load_smplx_npz

        smplx_model = load_smplx_npz(npz, device=g_body.device)
    return infer(g_body, g_face, smplx_model, None, config, args)


if __name__ == "__main__":
    main()
