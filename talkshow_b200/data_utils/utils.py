"""Audio front-end on the host (reference data_utils/utils.py:148-231, ``get_mfcc_ta``).

Body features: resample to 22 kHz, torchaudio MFCC(n_mfcc=64, n_fft=2048, n_mels=256,
hop=734 @30 fps / 1467 @15 fps, htk) -> (M, 64).  Face features: the raw 16 kHz waveform -> (N, 1).
torchaudio >= 2.9 needs torchcodec for ``torchaudio.load``; a scipy/wave reader is used instead
(decoding is not on the accelerated path — SURVEY.md §8f-1 lists the GPU MFCC as a "next" row).
"""
import numpy as np
import torch


def load_wav(path):
    """-> (float32 tensor [channels, samples] in [-1,1], sample_rate)."""
    from scipy.io import wavfile

    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 1:
        x = x[:, None]
    return torch.from_numpy(np.ascontiguousarray(x.T)), int(sr)


def mfcc_from_wave(audio, sr_0, sr=22000, fps=30):
    """audio [C,N] at sr_0 -> numpy (M, 64), exactly the transform chain of utils.py:150-177."""
    import torchaudio.transforms as ta_T

    if sr != sr_0:
        audio = ta_T.Resample(sr_0, sr)(audio)
    if audio.shape[0] > 1:
        audio = torch.mean(audio, dim=0, keepdim=True)
    hop_length = {15: 1467, 30: 734}[fps]
    mfcc = ta_T.MFCC(sample_rate=sr, n_mfcc=64,
                     melkwargs={"n_fft": 2048, "n_mels": 256, "hop_length": hop_length, "mel_scale": "htk"})
    return mfcc(audio).squeeze(dim=0).transpose(0, 1).numpy()


def get_mfcc_sepa(audio_fn, fps=15, sr=16000):
    """Reference data_utils/utils.py:234-263 (same signature): features of the first 2 s and of the remainder computed
    separately -> (numpy (M0+M1, 64), M0)."""
    import torchaudio.transforms as ta_T

    audio, sr_0 = load_wav(audio_fn)
    if sr != sr_0:
        audio = ta_T.Resample(sr_0, sr)(audio)
    if audio.shape[0] > 1:
        audio = torch.mean(audio, dim=0, keepdim=True)
    f0 = mfcc_from_wave(audio[:, :sr * 2], sr, sr=sr, fps=fps)
    f1 = mfcc_from_wave(audio[:, sr * 2:], sr, sr=sr, fps=fps)
    return np.concatenate((f0, f1), axis=0), f0.shape[0]


def get_mfcc_ta(audio_fn, eps=1e-6, fps=15, smlpx=False, sr=16000, n_mfcc=64, win_size=None, type="mfcc", am=None,
                am_sr=None, encoder_choice="mfcc"):
    """Same signature as the reference.  ``am`` given + encoder_choice='faceformer' -> raw 16 kHz wave
    (N,1) (utils.py:194-198, librosa.load(sr=16000) there); otherwise MFCC (M,64)."""
    audio, sr_0 = load_wav(audio_fn)
    if am is not None and encoder_choice == "faceformer":
        import torchaudio.transforms as ta_T

        mono = torch.mean(audio, dim=0, keepdim=True)
        if sr_0 != 16000:
            mono = ta_T.Resample(sr_0, 16000)(mono)
        return mono[0].numpy().reshape(-1, 1)
    if type != "mfcc":
        raise NotImplementedError("only type='mfcc' is on the TalkSHOW inference path")
    return mfcc_from_wave(audio, sr_0, sr=sr, fps=fps)
