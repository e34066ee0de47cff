"""Import the UNMODIFIED reference (yhw-yhw/TalkSHOW) from /root/reference in this
container so golden vectors can be generated from the reference's own code.

Only used by tests/golden/make_golden.py (run here, where /root/reference exists).
Nothing in the product, the -m gpu tests, smoke() or bench.py imports this.
Shim list follows SURVEY.md §8c / Appendix D.
"""
import importlib.machinery
import os
import sys
import types

REF = os.environ.get("TALKSHOW_REFERENCE", "/root/reference")


def import_reference():
    if "nets" in sys.modules and getattr(sys.modules["nets"], "__file__", "").startswith(REF):
        return sys.modules["nets"]
    sys.dont_write_bytecode = True
    from transformers import Wav2Vec2Config, Wav2Vec2Model, Wav2Vec2Processor  # noqa: F401  (before stubbing librosa)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m

    for n in ["librosa", "python_speech_features", "textgrid", "smplx", "matplotlib", "matplotlib.pyplot"]:
        if n not in sys.modules:
            stub(n)
    stub("torchaudio.sox_effects", apply_effects_tensor=None)
    os.chdir(REF)  # data_utils/mesh_dataset.py:16 opens a relative path at import
    sys.path.insert(0, REF)
    import nets.spg.wav2vec as w2v

    w2v.Wav2Vec2Model.from_pretrained = classmethod(
        lambda cls, name, *a, **k: cls(Wav2Vec2Config(attn_implementation="eager"))
    )
    import nets

    return nets
