"""Feature-side helpers of the reference's `preprocess.py` that sit directly either side of the hot path.

The numpy functions here keep the reference's names (including its spelling) and argument meaning so `train.py` /
`convert.py` callers can switch imports:

  coded_sps_normalization_fit_transoform / _transoform / _inverse_transoform   preprocess.py:106-133
  coded_sp_padding, wav_padding                                                preprocess.py:135-158
  logf0_statistics, pitch_conversion                                           preprocess.py:161-175
  transpose_in_list                                                            preprocess.py:63-68
  sample_train_data                                                            preprocess.py:207-238

WORLD analysis / synthesis (`world_decompose`, `world_encode_spectral_envelop`, ... preprocess.py:6-104) is CPU audio
code in pyworld + librosa and is NOT rebuilt: the wrappers below forward to pyworld when it is importable and raise a
clear ImportError otherwise (neither package exists in this image).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------------- WORLD (gated on pyworld)
def _pyworld():
    try:
        import pyworld
        return pyworld
    except ImportError as e:                                   # pragma: no cover - pyworld is absent in this image
        raise ImportError("WORLD analysis/synthesis needs the `pyworld` package (CPU audio code, outside the B200 hot path); "
                          "feed pre-extracted MCEP matrices instead (see cgvc.convert / cgvc.train)") from e


def world_decompose(wav, fs, frame_period=5.0):
    """f0 (harvest, 71..800 Hz), time axis, spectral envelope (cheaptrick), aperiodicity (d4c) -- preprocess.py:17-25."""
    pw = _pyworld()
    wav = np.asarray(wav, dtype=np.float64)
    f0, timeaxis = pw.harvest(wav, fs, frame_period=frame_period, f0_floor=71.0, f0_ceil=800.0)
    return f0, timeaxis, pw.cheaptrick(wav, f0, timeaxis, fs), pw.d4c(wav, f0, timeaxis, fs)


def world_encode_spectral_envelop(sp, fs, dim=24):
    return _pyworld().code_spectral_envelope(sp, fs, dim)


def world_decode_spectral_envelop(coded_sp, fs):
    pw = _pyworld()
    return pw.decode_spectral_envelope(coded_sp, fs, pw.get_cheaptrick_fft_size(fs))


def world_speech_synthesis(f0, decoded_sp, ap, fs, frame_period):
    return _pyworld().synthesize(f0, decoded_sp, ap, fs, frame_period).astype(np.float32)


def world_encode_data(wavs, fs, frame_period=5.0, coded_dim=24):
    cols = ([], [], [], [], [])
    for wav in wavs:
        f0, timeaxis, sp, ap = world_decompose(wav, fs, frame_period)
        for c, v in zip(cols, (f0, timeaxis, sp, ap, world_encode_spectral_envelop(sp, fs, coded_dim))):
            c.append(v)
    return cols


# --------------------------------------------------------------------------------------------- numpy feature helpers
def transpose_in_list(lst):
    return [np.asarray(a).T for a in lst]


def _fit(coded_sps):
    cat = np.concatenate(coded_sps, axis=1)
    return np.mean(cat, axis=1, keepdims=True), np.std(cat, axis=1, keepdims=True)


def coded_sps_normalization_fit_transoform(coded_sps):
    """Per-coefficient z-normalisation over all frames of all utterances ([24, frames] each).
    Returns (normalised list, mean [24,1], std [24,1])."""
    mean, std = _fit(coded_sps)
    return [(c - mean) / std for c in coded_sps], mean, std


def coded_sps_normalization_transoform(coded_sps, coded_sps_mean, coded_sps_std):
    return [(c - coded_sps_mean) / coded_sps_std for c in coded_sps]


def coded_sps_normalization_inverse_transoform(normalized_coded_sps, coded_sps_mean, coded_sps_std):
    return [c * coded_sps_std + coded_sps_mean for c in normalized_coded_sps]


def _split_pad(total):
    left = total // 2
    return left, total - left


def coded_sp_padding(coded_sp, multiple=4):
    """Zero-pad the frame axis of a [features, frames] matrix to a multiple of `multiple`, the smaller half in front."""
    frames = coded_sp.shape[1]
    left, right = _split_pad(-frames % multiple)
    return np.pad(coded_sp, ((0, 0), (left, right)), 'constant', constant_values=0)


def wav_padding(wav, sr, frame_period, multiple=4):
    """Pad a waveform so that WORLD yields a frame count that is a multiple of `multiple` (generator needs T % 4 == 0)."""
    assert wav.ndim == 1
    n = len(wav)
    hop = sr * frame_period / 1000
    padded = int((np.ceil((np.floor(n / hop) + 1) / multiple + 1) * multiple - 1) * hop)
    left, right = _split_pad(padded - n)
    return np.pad(wav, (left, right), 'constant', constant_values=0)


def logf0_statistics(f0s):
    """Mean / std of log f0 over voiced frames (unvoiced frames have f0 == 0 and are masked out)."""
    logs = np.ma.log(np.concatenate(f0s))
    return logs.mean(), logs.std()


def pitch_conversion(f0, mean_log_src, std_log_src, mean_log_target, std_log_target):
    """Log-Gaussian normalised pitch transformation; unvoiced frames (f0 == 0) map to 0 (exp(-inf))."""
    with np.errstate(divide='ignore'):
        z = (np.log(f0) - mean_log_src) / std_log_src
    return np.exp(z * std_log_target + mean_log_target)


def sample_train_data(dataset_A, dataset_B, n_frames=128):
    """One epoch's training pairs: both utterance index lists shuffled independently (global numpy RNG, seeded by the
    caller like train.py:13), truncated to the shorter list, one uniform random `n_frames` crop per utterance.
    Returns two arrays [num_samples, features, n_frames]."""
    num = min(len(dataset_A), len(dataset_B))
    order_A = np.arange(len(dataset_A)); order_B = np.arange(len(dataset_B))
    np.random.shuffle(order_A); np.random.shuffle(order_B)
    crops_A, crops_B = [], []
    for ia, ib in zip(order_A[:num], order_B[:num]):
        for utt, sink in ((dataset_A[ia], crops_A), (dataset_B[ib], crops_B)):     # A's crop is drawn before B's
            total = utt.shape[1]
            assert total >= n_frames
            s = np.random.randint(total - n_frames + 1)
            sink.append(utt[:, s:s + n_frames])
    return np.array(crops_A), np.array(crops_B)
