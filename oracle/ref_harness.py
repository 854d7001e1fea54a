"""TEST INFRASTRUCTURE — not part of the product path.

Loads the UNMODIFIED reference (rutgerswiselab/DCCF) from /root/reference/src and runs it on CPU so
that its outputs can pin the restatement in `oracle/dccf_oracle.py` and produce the golden fixtures
under `tests/golden/` (see `oracle/make_golden.py`).  Nothing is copied from the reference: its
modules are imported where they lie, under five environment shims:

  1. `pymining` (imported by src/utils/mining.py:5, unused on this path) is stubbed;
  2. `np.float`, `np.int`, `np.asfarray` (src/utils/utils.py:75-77, src/utils/rank_metrics.py:159)
     are restored for numpy >= 1.24;
  3. pandas >= 3 hands out read-only views; `DataProcessor.format_data_dict` results are copied so
     `np.random.shuffle` (src/utils/utils.py:91) can work in place;
  4. DCCF hard-codes CUDA (src/models/DCCF.py:55,64,72,87): `torch.cuda.current_device`,
     `Tensor.to('cuda:..')` and `torch.cuda.FloatTensor` are redirected to CPU tensors;
  5. `../result/` is never created by src/main.py:80-81,192.

The three random draws of DCCF.predict — confounder items (DCCF.py:72), feature noise (DCCF.py:87)
and the dropout mask (DCCF.py:94) — are recorded (or injected) through `RngTape`.

/root/reference does not exist on the GPU box: this module is only importable in the build
container; tests that need it are skipped elsewhere and rely on the committed fixtures.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_SRC = os.environ.get('DCCF_REFERENCE_SRC', '/root/reference/src')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_SRC, 'models', 'DCCF.py'))


class RngTape:
    """Record or replay the draws made inside DCCF.predict, one entry per predict call."""

    def __init__(self):
        self.calls = []       # list of dict(sample_item=LongTensor[P,S], noise=FloatTensor[N,F], masks=[FloatTensor[N,D]...])
        self.replay = None    # list of the same dicts to inject instead of drawing
        self._cur = None

    def begin_call(self):
        self._cur = {'sample_item': None, 'noise': None, 'masks': []}
        self.calls.append(self._cur)
        return self._cur

    def _injected(self, key, idx=None):
        if self.replay is None:
            return None
        ent = self.replay[len(self.calls) - 1]
        val = ent[key]
        if idx is not None:
            val = val[idx]
        return val


_TAPE = RngTape()
_STATE = {'loaded': None, 'device_rng': None}


def use_device_rng(generator):
    """On a GPU the reference draws the feature noise (DCCF.py:87) and the dropout masks (DCCF.py:94) from the CUDA
    generator: the torch CPU generator only sees the confounder draws (DCCF.py:72).  With a private `generator` here the
    CPU harness consumes the CPU generator the way a GPU run of the reference does; None (default) = everything on the
    CPU generator, as a literal `--gpu ''` run would."""
    _STATE['device_rng'] = generator


def tape():
    return _TAPE


def reset_tape(replay=None):
    _TAPE.calls = []
    _TAPE.replay = replay
    _TAPE._cur = None


class _NoiseBuffer:
    """Stands in for `torch.cuda.FloatTensor(shape)`: only `.normal_(std=)` is used (DCCF.py:87)."""

    def __init__(self, shape):
        self.shape = tuple(shape)

    def normal_(self, mean=0.0, std=1.0):
        inj = _TAPE._injected('noise')
        if inj is not None:
            out = inj.clone()
        else:
            out = torch.empty(self.shape, dtype=torch.float32).normal_(mean=mean, std=std, generator=_STATE['device_rng'])
        if _TAPE._cur is not None:
            _TAPE._cur['noise'] = out.clone()
        return out


class _TapedDropout(torch.nn.Module):
    """`torch.nn.Dropout(p)(x)` with the mask exposed: x * bernoulli(1-p)/(1-p), always in training
    mode exactly like the freshly constructed module at DCCF.py:94; identity when p == 0."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if self.p == 0.0:
            return x
        cur = _TAPE._cur
        idx = len(cur['masks']) if cur is not None else 0
        inj = _TAPE._injected('masks', idx)
        if inj is not None:
            mask = inj.clone()
        else:
            mask = torch.empty_like(x).bernoulli_(1.0 - self.p, generator=_STATE['device_rng']).div_(1.0 - self.p)
        if cur is not None:
            cur['masks'].append(mask.clone())
        return x * mask


def load_reference():
    """Import the reference modules under the shims; returns a namespace with the classes."""
    if _STATE['loaded'] is not None:
        return _STATE['loaded']
    if not reference_available():
        raise RuntimeError('reference sources not found at %s' % REFERENCE_SRC)

    # shim 1
    if 'pymining' not in sys.modules:
        sys.modules['pymining'] = types.SimpleNamespace(itemmining=None, assocrules=None, perftesting=None)
    # shim 2
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(np, 'int'):
        np.int = int
    if not hasattr(np, 'asfarray'):
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)

    # the reference uses top-level package names (models, utils, runners ...): import them from its tree
    # without leaving its directory on sys.path for good
    clash = [m for m in ('utils', 'models', 'runners', 'data_loaders', 'data_processor') if m in sys.modules]
    if clash:
        raise RuntimeError('modules %s already imported; the reference needs these top-level names' % clash)
    sys.path.insert(0, REFERENCE_SRC)
    try:
        from utils import utils as ref_utils, global_p as ref_global_p, rank_metrics as ref_rank_metrics
        from data_loaders.DataLoader import DataLoader as RefDataLoader
        from data_processor.DataProcessor import DataProcessor as RefDataProcessor
        from models.BaseModel import BaseModel as RefBaseModel
        from models.DCCF import DCCF as RefDCCF
        from models.IPSBiasedMF import IPSBiasedMF as RefIPSBiasedMF
        from runners.BaseRunner import BaseRunner as RefBaseRunner
        import models.DCCF as ref_dccf_module
    finally:
        sys.path.remove(REFERENCE_SRC)

    # shim 3
    orig_format = RefDataProcessor.format_data_dict

    def format_data_dict(self, df):
        data = orig_format(self, df)
        for k in ('uid', 'iid'):
            if k in data:
                data[k] = np.array(data[k], copy=True)
        return data

    RefDataProcessor.format_data_dict = format_data_dict

    # predict wrapper: opens a tape entry per call and records the confounder draw
    orig_predict = RefDCCF.predict

    def predict(self, feed_dict):
        _TAPE.begin_call()
        return orig_predict(self, feed_dict)

    RefDCCF.predict = predict

    ns = types.SimpleNamespace(
        utils=ref_utils, global_p=ref_global_p, rank_metrics=ref_rank_metrics, DataLoader=RefDataLoader,
        DataProcessor=RefDataProcessor, BaseModel=RefBaseModel, DCCF=RefDCCF, IPSBiasedMF=RefIPSBiasedMF,
        BaseRunner=RefBaseRunner, dccf_module=ref_dccf_module)
    _STATE['loaded'] = ns
    return ns


@contextlib.contextmanager
def cpu_shims():
    """Shim 4 + the RNG taps, active only inside the `with` block."""
    saved = {
        'current_device': torch.cuda.current_device,
        'to': torch.Tensor.to,
        'FloatTensor': getattr(torch.cuda, 'FloatTensor', None),
        'Dropout': torch.nn.Dropout,
        'randint': torch.randint,
        'device_count': torch.cuda.device_count,
    }
    orig_to = torch.Tensor.to
    orig_randint = torch.randint

    def to(self, *args, **kwargs):
        if args and isinstance(args[0], str) and args[0].startswith('cuda'):
            args = ('cpu',) + tuple(args[1:])
        return orig_to(self, *args, **kwargs)

    def randint(*args, **kwargs):
        inj = _TAPE._injected('sample_item') if _TAPE._cur is not None and _TAPE._cur['sample_item'] is None else None
        out = inj.clone() if inj is not None else orig_randint(*args, **kwargs)
        if _TAPE._cur is not None and _TAPE._cur['sample_item'] is None:
            _TAPE._cur['sample_item'] = out.clone()
        return out

    torch.cuda.current_device = lambda: 0
    torch.cuda.device_count = lambda: 0
    torch.Tensor.to = to
    torch.cuda.FloatTensor = _NoiseBuffer
    torch.nn.Dropout = _TapedDropout
    torch.randint = randint
    try:
        yield
    finally:
        torch.cuda.current_device = saved['current_device']
        torch.cuda.device_count = saved['device_count']
        torch.Tensor.to = saved['to']
        if saved['FloatTensor'] is not None:
            torch.cuda.FloatTensor = saved['FloatTensor']
        torch.nn.Dropout = saved['Dropout']
        torch.randint = saved['randint']


def build_reference_model(ref, data_dir, dataset, sentence_model, user_num, item_num, sample_num=10, attribute_num=2,
                          std=0.1, n_layers=1, random_seed=2019, dim=64, model_path='/tmp/dccf_ref_model.pt'):
    """Construct the reference DCCF exactly as src/main.py:137-150 does (ctor + apply(init_paras))."""
    model = ref.DCCF(path=data_dir, dataset=dataset, sentence_model=sentence_model, sample_num=sample_num,
                     attribute_num=attribute_num, std=std, label_min=0, label_max=1, feature_num=0,
                     user_num=user_num, item_num=item_num, u_vector_size=dim, i_vector_size=dim, n_layers=n_layers,
                     random_seed=random_seed, model_path=model_path)
    model.apply(model.init_paras)
    return model
