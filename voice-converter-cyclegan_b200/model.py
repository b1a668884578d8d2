"""`CycleGAN`: the reference's model class (model.py:7-169 of /root/reference) on the native B200 engine.

Same constructor and method signatures as the reference; the TensorFlow-1 session is replaced by libcgvc.so
(hand-written sm_100a CUDA behind the C ABI in include/cgvc.h).  PyTorch is used for device storage
(arenas, staging) only -- no torch op runs on the hot path.

Differences a user of the reference should know:
  * weights are glorot-uniform like TF's default, drawn from `seed`
  * minibatches larger than 1 are first-class (the reference hard-codes 1, train.py:16)
  * `train`/`test` accept host numpy arrays (as in the reference) or CUDA torch tensors (zero-copy)
  * checkpoints are .npz files keyed by the TF variable names (plus `<var>/Adam`, `<var>/Adam_1`, `adam_step`)
  * with `torch.distributed` initialised and `data_parallel=True`, one process per GPU trains data-parallel:
    a single NCCL all-reduce of the flat gradient arena per step
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from datetime import datetime

import numpy as np
import torch

from . import _native as N
from .module import discriminator as _discriminator, generator_gatedcnn as _generator_gatedcnn


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class CycleGAN(object):

    def __init__(self, num_features, discriminator=_discriminator, generator=_generator_gatedcnn, mode='train',
                 log_dir='./log', *, max_batch=1, max_frames=None, precision='bf16x3', device=None, seed=0,
                 data_parallel=False, summary_interval=0):
        for net in (discriminator, generator):
            if not hasattr(net, "check_engine_table"):
                raise TypeError("CycleGAN(discriminator=..., generator=...) takes network descriptors (cgvc.module.generator_gatedcnn / "
                                "cgvc.module.discriminator or equivalents), not %r: the native engine runs its own kernel graph" % (net,))
        if not torch.cuda.is_available():
            raise RuntimeError("CycleGAN needs a CUDA device (sm_100a); there is no CPU fallback")
        self.num_features = num_features
        self.input_shape = [None, num_features, None]
        self.discriminator = discriminator
        self.generator = generator
        self.mode = mode
        self.precision = precision
        self._lib = N.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self._handle = C.c_void_p(0)
        self._max_batch = int(max_batch)
        self._max_frames = int(max_frames) if max_frames else 128
        self._arenas = {}
        self._options = {}
        self._create_engine()
        # the descriptors state the architecture the caller expects (model.py:14-15); the engine must implement exactly that
        for scope in ("generator_A2B", "generator_B2A"):
            generator.check_engine_table(self._table, scope, num_features)
        for scope in ("discriminator_A", "discriminator_B"):
            discriminator.check_engine_table(self._table, scope, num_features)
        self._init_params(seed)
        self._rank, self._nranks = 0, 1
        self._data_parallel = bool(data_parallel)
        if data_parallel:
            self._attach_communicator()
        self.train_step = 0
        self.last_losses = None
        self.writer = None
        self.summary_interval = summary_interval
        if self.mode == 'train':
            now = datetime.now()
            self.log_dir = os.path.join(log_dir, now.strftime('%Y%m%d-%H%M%S'))
            if summary_interval:
                self.generator_summaries, self.discriminator_summaries = self.summary()

    # ------------------------------------------------------------------ engine plumbing
    def _chk(self, code):
        N.check(self._handle, code)

    def _create_engine(self):
        cfg = N.Config(self.num_features, self._max_batch, self._max_frames, N.PRECISIONS[self.precision],
                       self.device.index, 1 if self.mode == 'train' else 0)
        h = C.c_void_p(0)
        code = self._lib.cgvc_create(C.byref(cfg), C.byref(h))
        if code != 0:
            raise N.CgvcError(code, (self._lib.cgvc_last_error(None) or b"?").decode())
        self._handle = h
        nt, ne = C.c_int(0), C.c_size_t(0)
        self._chk(self._lib.cgvc_param_count(h, C.byref(nt), C.byref(ne)))
        self.n_params = ne.value
        self._table = OrderedDict()
        for i in range(nt.value):
            name, off, nd, shp = C.c_char_p(), C.c_size_t(), C.c_int(), (C.c_int * 4)()
            self._chk(self._lib.cgvc_param_info(h, i, C.byref(name), C.byref(off), C.byref(nd), C.byref(shp)))
            self._table[name.value.decode()] = (off.value, tuple(shp[k] for k in range(nd.value)))
        self._generator_end = max(o + int(np.prod(s)) for n, (o, s) in self._table.items() if 'generator' in n)
        kinds = [N.ARENA_PARAM, N.ARENA_WORK]
        if self.mode == 'train':
            kinds += [N.ARENA_GRAD, N.ARENA_ADAM_M, N.ARENA_ADAM_V]
        for kind in kinds:
            nbytes = C.c_size_t(0)
            self._chk(self._lib.cgvc_arena_bytes(h, kind, C.byref(nbytes)))
            old = self._arenas.get(kind)
            if kind == N.ARENA_WORK or old is None:
                t = torch.empty((nbytes.value + 3) // 4, dtype=torch.float32, device=self.device)
                if kind != N.ARENA_WORK:
                    t.zero_()
                self._arenas[kind] = t
            t = self._arenas[kind]
            self._chk(self._lib.cgvc_bind_arena(h, kind, _ptr(t), t.numel() * 4))
        self._losses = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._losses_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._staging = {}
        for name, value in self._options.items():            # the engine is re-created when batch / frames outgrow it
            self._chk(self._lib.cgvc_set_option(self._handle, name.encode(), int(value)))

    def set_option(self, name, value):
        """Engine options of include/cgvc.h (`two_streams`, `cuda_graph`, `fuse_in`, `fuse_bwd`, `debug_taps`); remembered across
        engine re-creations."""
        self._chk(self._lib.cgvc_set_option(self._handle, name.encode(), int(value)))
        self._options[name] = int(value)

    def _ensure_capacity(self, batch, frames):
        if batch <= self._max_batch and frames <= self._max_frames:
            return
        step = C.c_longlong(0)
        self._lib.cgvc_get_adam_step(self._handle, C.byref(step))
        self._lib.cgvc_destroy(self._handle)
        self._max_batch = max(batch, self._max_batch)
        self._max_frames = max(frames, self._max_frames)
        self._arenas.pop(N.ARENA_WORK, None)
        torch.cuda.empty_cache()
        self._create_engine()
        self._lib.cgvc_set_adam_step(self._handle, step)
        if self._data_parallel:
            # cgvc_destroy freed the NCCL communicator with the old engine: a data-parallel model must get a new one, or it would
            # silently train without the all-reduce.  Collective: every rank has to grow in the same call (same batch / frames).
            self._attach_communicator()
        else:
            self._params_updated()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _params_updated(self):
        self._chk(self._lib.cgvc_params_updated(self._handle, self._stream()))

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self._lib.cgvc_destroy(self._handle)
                self._handle = C.c_void_p(0)
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def param_names(self):
        return list(self._table.keys())

    def _view(self, arena, name):
        off, shape = self._table[name]
        n = int(np.prod(shape))
        return self._arenas[arena][off:off + n].view(shape)

    def _init_params(self, seed):
        """glorot-uniform kernels, zero biases / beta, unit gamma (tf.get_variable defaults, Appendix A.3)."""
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(seed))
        for name, (off, shape) in self._table.items():
            v = self._view(N.ARENA_PARAM, name)
            if name.endswith("/kernel"):
                rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
                fan_in, fan_out = rf * shape[-2], rf * shape[-1]
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                v.copy_((torch.rand(shape, generator=gen, device=self.device) * 2 - 1) * lim)
            elif name.endswith("/gamma"):
                v.fill_(1.0)
            else:
                v.zero_()
        self._params_updated()

    def get_params(self):
        """name -> numpy array (TF layouts)."""
        torch.cuda.synchronize(self.device)
        return OrderedDict((n, self._view(N.ARENA_PARAM, n).cpu().numpy()) for n in self._table)

    def set_params(self, params):
        """Inject weights (name -> array-like in TF layout); missing names keep their value."""
        for n, a in params.items():
            if n not in self._table:
                raise KeyError("unknown variable %r" % n)
            t = torch.as_tensor(np.asarray(a, dtype=np.float32))
            self._view(N.ARENA_PARAM, n).copy_(t.reshape(self._table[n][1]))
        self._params_updated()

    def get_grads(self):
        torch.cuda.synchronize(self.device)
        return OrderedDict((n, self._view(N.ARENA_GRAD, n).cpu().numpy()) for n in self._table)

    # ------------------------------------------------------------------ data parallel
    def _attach_communicator(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("data_parallel=True needs torch.distributed to be initialised (one process per GPU)")
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = (C.c_char * 128)()
        if rank == 0:
            self._chk(self._lib.cgvc_comm_unique_id(self._handle, idbuf))
        obj = [bytes(idbuf)]
        dist.broadcast_object_list(obj, src=0)
        idbuf = (C.c_char * 128).from_buffer_copy(obj[0])
        self._chk(self._lib.cgvc_comm_init(self._handle, idbuf, rank, world))
        self._rank, self._nranks = rank, world
        # every replica must start from identical weights: broadcast rank 0's parameter arena
        dist.broadcast(self._arenas[N.ARENA_PARAM], src=0)
        self._params_updated()

    # ------------------------------------------------------------------ staging
    def _to_device(self, x, key):
        """Host array (any float dtype, model.py feeds float64) or CUDA tensor -> fp32 CUDA tensor [B,F,T]."""
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.to(dtype=torch.float32).contiguous()
        a = np.asarray(x)
        if a.ndim != 3 or a.shape[1] != self.num_features:
            raise ValueError("expected [batch, %d, frames], got %r" % (self.num_features, a.shape))
        st = self._staging.get(key)
        if st is None or st[0].shape != a.shape:
            st = (torch.empty(a.shape, dtype=torch.float32).pin_memory(),
                  torch.empty(a.shape, dtype=torch.float32, device=self.device))
            self._staging[key] = st
        st[0].numpy()[...] = a            # cast to fp32 at the boundary, like the placeholder feed (model.py:35-42)
        st[1].copy_(st[0], non_blocking=True)
        return st[1]

    # ------------------------------------------------------------------ the reference API
    def train(self, input_A, input_B, lambda_cycle, lambda_identity, generator_learning_rate, discriminator_learning_rate):
        """One G step + one D step from the same pre-update weights (model.py:110-125).
        Returns (generator_loss, discriminator_loss) as fp32 scalars (pre-update values)."""
        A = self._to_device(input_A, "A")
        B = self._to_device(input_B, "B")
        if A.shape != B.shape:
            raise ValueError("input_A and input_B must have the same shape")
        batch, _, frames = A.shape
        self._ensure_capacity(batch, frames)
        self._chk(self._lib.cgvc_train_step(self._handle, _ptr(A), _ptr(B), batch, frames,
                                            float(lambda_cycle), float(lambda_identity),
                                            float(generator_learning_rate), float(discriminator_learning_rate),
                                            None, None, _ptr(self._losses), self._stream()))
        self._losses_host.copy_(self._losses, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        l = self._losses_host.numpy()
        self.last_losses = {k: float(v) for k, v in zip(N.LOSS_NAMES, l)}
        if self.writer is not None and self.summary_interval and self.train_step % self.summary_interval == 0:
            self._write_summaries()
        self.train_step += 1
        return np.float32(l[4]), np.float32(l[7])

    def train_async(self, A_dev, B_dev, lambda_cycle, lambda_identity, generator_learning_rate, discriminator_learning_rate):
        """Device-resident variant: enqueue one step, no host synchronisation.  Losses land in self._losses."""
        batch, _, frames = A_dev.shape
        self._chk(self._lib.cgvc_train_step(self._handle, _ptr(A_dev), _ptr(B_dev), batch, frames,
                                            float(lambda_cycle), float(lambda_identity),
                                            float(generator_learning_rate), float(discriminator_learning_rate),
                                            None, None, _ptr(self._losses), self._stream()))
        self.train_step += 1

    def fetch_losses(self):
        """(generator_loss, discriminator_loss) of the most recent train_async step: the one device -> host read (32 bytes) and stream
        synchronisation a device-resident training loop needs, at the steps it logs."""
        self._losses_host.copy_(self._losses, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        l = self._losses_host.numpy()
        self.last_losses = {k: float(v) for k, v in zip(N.LOSS_NAMES, l)}
        if self.writer is not None and self.summary_interval:
            self._write_summaries()
        return np.float32(l[4]), np.float32(l[7])

    def compute_gradients(self, input_A, input_B, lambda_cycle, lambda_identity):
        """Forward + backward only (what the two `minimize` calls differentiate, model.py:107-108).
        Returns (losses dict, generation_A, generation_B); gradients via get_grads()."""
        A = self._to_device(input_A, "A"); B = self._to_device(input_B, "B")
        batch, _, frames = A.shape
        self._ensure_capacity(batch, frames)
        gA = torch.empty_like(A); gB = torch.empty_like(B)
        self._chk(self._lib.cgvc_compute_gradients(self._handle, _ptr(A), _ptr(B), batch, frames, float(lambda_cycle),
                                                   float(lambda_identity), _ptr(gA), _ptr(gB), _ptr(self._losses), self._stream()))
        torch.cuda.synchronize(self.device)
        l = self._losses.cpu().numpy()
        return {k: float(v) for k, v in zip(N.LOSS_NAMES, l)}, gA.cpu().numpy(), gB.cpu().numpy()

    def test(self, inputs, direction):
        """Generator forward (model.py:128-137).  inputs [B, 24, T] with T a multiple of 4."""
        if direction == 'A2B':
            d = 0
        elif direction == 'B2A':
            d = 1
        else:
            raise Exception('Conversion direction must be specified.')
        x = self._to_device(inputs, "test")
        batch, _, frames = x.shape
        self._ensure_capacity(batch, frames)
        y = torch.empty_like(x)
        self._chk(self._lib.cgvc_generator_forward(self._handle, d, _ptr(x), _ptr(y), batch, frames, self._stream()))
        if isinstance(inputs, torch.Tensor) and inputs.is_cuda:
            return y
        out = torch.empty(y.shape, dtype=torch.float32).pin_memory()
        out.copy_(y, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out.numpy().copy()

    def discriminate(self, inputs, which):
        """Discriminator forward (module.py:188-213): which in {'A','B'}; returns [B, 6, T/16, 1]."""
        x = self._to_device(inputs, "disc")
        batch, _, frames = x.shape
        self._ensure_capacity(batch, frames)
        y = torch.empty((batch, self.num_features // 4, frames // 16, 1), dtype=torch.float32, device=self.device)
        self._chk(self._lib.cgvc_discriminator_forward(self._handle, {'A': 0, 'B': 1}[which], _ptr(x), _ptr(y), batch, frames, self._stream()))
        torch.cuda.synchronize(self.device)
        return y.cpu().numpy()

    def set_debug_taps(self, on=True):
        """Keep the fp32 copy of every generator layer output of the next test() calls for debug_activation() (parity tests).
        Off by default: the conversion path then writes only what the next layer reads."""
        self.set_option("debug_taps", 1 if on else 0)

    def debug_activation(self, name):
        n = C.c_size_t(0)
        self._chk(self._lib.cgvc_debug_activation(self._handle, name.encode(), None, 0, C.byref(n), self._stream()))
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        self._chk(self._lib.cgvc_debug_activation(self._handle, name.encode(), _ptr(out), n.value, C.byref(n), self._stream()))
        torch.cuda.synchronize(self.device)
        return out.cpu().numpy()

    def save(self, directory, filename):
        """model.py:140-146.  Writes <directory>/<filename>.npz keyed by TF variable names; returns the joined path."""
        if not os.path.exists(directory):
            os.makedirs(directory)
        path = os.path.join(directory, filename)
        torch.cuda.synchronize(self.device)
        blob = {}
        for n in self._table:
            blob[n] = self._view(N.ARENA_PARAM, n).cpu().numpy()
            if self.mode == 'train':
                blob[n + "/Adam"] = self._view(N.ARENA_ADAM_M, n).cpu().numpy()
                blob[n + "/Adam_1"] = self._view(N.ARENA_ADAM_V, n).cpu().numpy()
        step = C.c_longlong(0)
        self._lib.cgvc_get_adam_step(self._handle, C.byref(step))
        blob["adam_step"] = np.int64(step.value)
        blob["train_step"] = np.int64(self.train_step)
        with open(path + ".npz" if not path.endswith(".npz") else path, "wb") as f:
            np.savez(f, **blob)
        return path

    def load(self, filepath):
        """model.py:148-150.  Accepts this engine's `.npz` checkpoints and TensorFlow V2 bundles written by the reference's
        `tf.train.Saver` (`<filepath>.index` + `<filepath>.data-*`, e.g. the SF1-TM1 model the reference's README publishes):
        the variable names are the same in both."""
        from . import tf_checkpoint as tfc
        if tfc.is_bundle(filepath):
            return self._load_tf_bundle(filepath)
        p = filepath if filepath.endswith(".npz") else filepath + ".npz"
        z = np.load(p)
        for n in self._table:
            self._view(N.ARENA_PARAM, n).copy_(torch.from_numpy(z[n]))
            if self.mode == 'train' and (n + "/Adam") in z:
                self._view(N.ARENA_ADAM_M, n).copy_(torch.from_numpy(z[n + "/Adam"]))
                self._view(N.ARENA_ADAM_V, n).copy_(torch.from_numpy(z[n + "/Adam_1"]))
        if "adam_step" in z:
            self._lib.cgvc_set_adam_step(self._handle, int(z["adam_step"]))
        if "train_step" in z:
            self.train_step = int(z["train_step"])
        self._params_updated()

    def _load_tf_bundle(self, prefix):
        from . import tf_checkpoint as tfc
        want = set(self._table)
        if self.mode == 'train':
            want |= {n + s for n in self._table for s in ("/Adam", "/Adam_1")} | {"beta2_power", "beta2_power_1"}
        t = tfc.read_checkpoint(prefix, names=want)
        missing = [n for n in self._table if n not in t]
        if missing:
            raise KeyError("TensorFlow checkpoint %s lacks %d variables, e.g. %s" % (prefix, len(missing), missing[0]))
        for n, (off, shape) in self._table.items():
            a = np.asarray(t[n], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError("%s: checkpoint shape %r, expected %r" % (n, a.shape, shape))
            self._view(N.ARENA_PARAM, n).copy_(torch.from_numpy(a))
            if self.mode == 'train' and (n + "/Adam") in t and (n + "/Adam_1") in t:
                self._view(N.ARENA_ADAM_M, n).copy_(torch.from_numpy(np.asarray(t[n + "/Adam"], dtype=np.float32)))
                self._view(N.ARENA_ADAM_V, n).copy_(torch.from_numpy(np.asarray(t[n + "/Adam_1"], dtype=np.float32)))
        if self.mode == 'train' and "beta2_power" in t:
            # tf.train.AdamOptimizer keeps beta2^t (Appendix A.6): recover the step count both optimizers share
            b2p = float(np.asarray(t["beta2_power"]).reshape(-1)[0])
            if 0.0 < b2p < 1.0:
                self._lib.cgvc_set_adam_step(self._handle, int(round(math.log(b2p) / math.log(0.999))))
            elif b2p <= 0.0:
                # beta2^t underflows fp32 after ~87k steps: any large t gives the same (unit) bias correction
                self._lib.cgvc_set_adam_step(self._handle, 1000000)
        self._params_updated()

    def summary(self):
        """model.py:153-169: the same 8 scalar tags, written every `summary_interval` steps."""
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(self.log_dir)
        except Exception:
            self.writer = None
        g = ['generator_summaries/' + n for n in N.LOSS_NAMES[:5]]
        d = ['discriminator_summaries/' + n for n in N.LOSS_NAMES[5:]]
        return g, d

    def _write_summaries(self):
        for n, v in self.last_losses.items():
            scope = 'generator_summaries/' if not n.startswith('discriminator') else 'discriminator_summaries/'
            self.writer.add_scalar(scope + n, v, self.train_step)
