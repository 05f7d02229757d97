"""Record every call the reference's OWN transition unit tests make into its transition modules.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_kat.py

The reference's tests (tests/free_range_zoo/envs/*/env/transitions/test_*.py) build small hand-written states, call a
transition ``nn.Module`` with a fixed "randomness" grid and compare the result with literal expected tensors.  Here
the modules' ``forward`` is wrapped so that, while those tests run (and pass) under the import shim, each call's
module buffers, inputs and outputs are captured.  The captured records -- known-answer vectors authored by the
reference's maintainers -- are written to tests/golden/kat_<domain>.npz and replayed against the oracle's transition
functions by tests/test_oracle_kat.py.
"""
import dataclasses
import inspect
import json
import os
import sys
import unittest

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402

ref_shim.install()
os.chdir(ref_shim.REFERENCE_ROOT)

DOMAINS = {
    'wildfire': 'free_range_zoo.envs.wildfire.env.transitions',
    'rideshare': 'free_range_zoo.envs.rideshare.env.transitions',
    'cybersecurity': 'free_range_zoo.envs.cybersecurity.env.transitions',
}


def flatten(prefix, value, out):
    if isinstance(value, torch.Tensor):
        out[prefix] = value.detach().cpu().numpy().copy()
    elif dataclasses.is_dataclass(value):
        for field in dataclasses.fields(value):
            flatten(f'{prefix}.{field.name}', getattr(value, field.name), out)
    elif isinstance(value, (tuple, list)):
        for i, item in enumerate(value):
            flatten(f'{prefix}.{i}', item, out)
    elif isinstance(value, (bool, int, float)):
        out[prefix] = np.asarray(value)


def record_domain(domain, package):
    import importlib
    import pkgutil
    records = []
    module = importlib.import_module(package)
    classes = []
    for info in pkgutil.iter_modules(module.__path__):
        sub = importlib.import_module(f'{package}.{info.name}')
        for name, cls in vars(sub).items():
            if isinstance(cls, type) and issubclass(cls, torch.nn.Module) and cls.__module__ == sub.__name__:
                classes.append(cls)

    def wrap(cls):
        original = cls.forward
        signature = inspect.signature(original)

        def forward(self, *args, **kwargs):
            bound = signature.bind(self, *args, **kwargs)
            bound.apply_defaults()
            record = {'transition': cls.__name__, 'test': current_test[0]}
            arrays = {}
            for name, value in bound.arguments.items():
                if name != 'self':
                    flatten(f'in.{name}', value, arrays)
            for name, buffer in list(self.named_buffers()) + list(self.named_parameters()):
                arrays[f'buf.{name}'] = buffer.detach().cpu().numpy().copy()
            for name, value in vars(self).items():
                if isinstance(value, (bool, int, float)) and not name.startswith('_') and name != 'training':
                    arrays[f'buf.{name}'] = np.asarray(value)
            result = original(self, *args, **kwargs)
            flatten('out', result, arrays)
            record['arrays'] = arrays
            records.append(record)
            return result

        cls.forward = forward

    current_test = [None]
    for cls in classes:
        wrap(cls)

    class Tracker(unittest.TextTestResult):

        def startTest(self, test):
            current_test[0] = test.id()
            super().startTest(test)

    suite = unittest.defaultTestLoader.discover(f'tests/free_range_zoo/envs/{domain}/env/transitions', top_level_dir='.')
    runner = unittest.TextTestRunner(resultclass=Tracker, verbosity=0, stream=open(os.devnull, 'w'))
    outcome = runner.run(suite)
    assert outcome.wasSuccessful(), (domain, outcome.failures, outcome.errors)

    data, index = {}, []
    for i, record in enumerate(records):
        index.append({'transition': record['transition'], 'test': record['test'], 'keys': sorted(record['arrays'])})
        for key, array in record['arrays'].items():
            data[f'r{i}/{key}'] = array
    data['index'] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f'kat_{domain}.npz')
    np.savez_compressed(path, **data)
    print(f'{domain}: {outcome.testsRun} reference tests run ({len(outcome.skipped)} skipped), {len(records)} transition calls '
          f'recorded, {os.path.getsize(path) / 1024:.1f} KiB')


if __name__ == '__main__':
    for domain, package in DOMAINS.items():
        record_domain(domain, package)
