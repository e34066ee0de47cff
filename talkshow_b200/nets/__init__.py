"""Drop-in mirror of the reference's ``nets`` package surface (nets/__init__.py:1-3,
nets/init_model.py): the three inference wrappers with the reference's constructor,
``load_state_dict`` and ``infer_on_audio`` / ``generate`` signatures, running on the CUDA engine."""
from .smplx_body_pixel import TrainWrapper as s2g_body_pixel  # noqa: F401
from .smplx_body_vq import TrainWrapper as s2g_body_vq  # noqa: F401
from .smplx_face import TrainWrapper as s2g_face  # noqa: F401


def init_model(model_name, args, config):
    """nets/init_model.py: name -> wrapper (unknown names raise NotImplementedError like demo.py:52)."""
    table = {"s2g_face": s2g_face, "s2g_body_vq": s2g_body_vq, "s2g_body_pixel": s2g_body_pixel}
    if model_name not in table:
        raise NotImplementedError(model_name)
    return table[model_name](args, config)
