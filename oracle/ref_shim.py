"""Import shim that lets the *unmodified* reference run on CPU -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Used by ``tests/golden/gen_golden.py`` (fixture generation, from ``/root/reference`` in the build container) and by
``bench.py --impl reference`` / the ``cpu_baseline`` leg (from ``oracle/_ref/``, the byte-for-byte copy of the
reference package that ``oracle/build_ref.py`` makes; it is git-ignored but travels to the GPU box).  The product
package never imports this module.

The reference imports seven third-party modules that are not installed here and cannot be installed (no
network): tensordict, gymnasium, pettingzoo, free_range_rust, supersuit, sqlalchemy, pygame.  None of them carries
transition arithmetic; they are replaced by inert stand-ins so the reference's real torch
code (``free_range_zoo/envs/*/env/*.py``) executes unchanged.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """Directory that holds the reference's ``free_range_zoo`` package: the checkout in the build container, else the
    copy under ``oracle/_ref`` (oracle/build_ref.py); None when neither exists."""
    override = os.environ.get('FRZ_REFERENCE_ROOT')  # (tests of the copy itself)
    for root in ([override] if override else []) + ['/root/reference', os.path.join(_HERE, '_ref')]:
        if os.path.isfile(os.path.join(root, 'free_range_zoo', '__init__.py')):
            return root
    return None


REFERENCE_ROOT = reference_root()


def _module(name):
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


class _TensorDict(dict):
    """dict with the two attributes the reference reads (wildfire.py:709-717)."""

    def __init__(self, source=None, batch_size=None, device=None, **_):
        super().__init__(source or {})
        self.batch_size = batch_size
        self.device = device


class _AgentSelector:
    """pettingzoo.utils.agent_selector contract used by utils/env.py:155-156,220,242."""

    def __init__(self, order):
        self.reinit(order)

    def reinit(self, order):
        self.agent_order = list(order)
        self._current = 0
        self.selected_agent = None

    def reset(self):
        self.reinit(self.agent_order)
        return self.next()

    def next(self):
        self._current = (self._current + 1) % len(self.agent_order)
        self.selected_agent = self.agent_order[self._current - 1]
        return self.selected_agent

    def is_last(self):
        return self.selected_agent == self.agent_order[-1]

    def is_first(self):
        return self.selected_agent == self.agent_order[0]


class _AECEnv:

    def __init__(self, *args, **kwargs):
        pass

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def unwrapped(self):
        return self


class _PassThroughWrapper:
    """OrderEnforcingWrapper / BaseWrapper stand-in: forwards everything to the wrapped env."""

    def __init__(self, env):
        object.__setattr__(self, 'env', env)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, 'env'), name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class _AecToParallel:

    def __init__(self, aec_env):
        self.aec_env = aec_env

    def __getattr__(self, name):
        return getattr(self.aec_env, name)

    def action_space(self, agent):
        return self.aec_env.action_space(agent)

    def observation_space(self, agent):
        return self.aec_env.observation_space(agent)


class _RecordSpace:
    """Record-only stand-in for free_range_rust.Space (no sampling arithmetic is pinned by the reference)."""

    def __init__(self, kind, *args, **kwargs):
        self.kind, self.args, self.kwargs = kind, args, kwargs

    def __eq__(self, other):
        return (self.kind, self.args, self.kwargs) == (other.kind, other.args, other.kwargs)

    def __hash__(self):
        return hash((self.kind, repr(self.args), repr(self.kwargs)))

    def __repr__(self):
        return f'{self.kind}{self.args}{self.kwargs}'


class _SpaceFactory:

    def __getattr__(self, kind):
        return lambda *a, **k: _RecordSpace(kind, *a, **k)


def install():
    """Seed sys.modules with the stand-ins and put the reference on sys.path. Idempotent."""
    if 'free_range_zoo' in sys.modules:
        return
    if REFERENCE_ROOT is None:
        raise ImportError('the reference package is neither at /root/reference nor under oracle/_ref '
                          '(run oracle/build_ref.py where /root/reference exists)')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    td = _module('tensordict')
    td.TensorDict = _TensorDict
    _module('tensordict.tensordict').TensorDict = _TensorDict

    gym = _module('gymnasium')
    gym.Space = type('Space', (), {})
    gym.Env = type('Env', (), {})
    gym.Wrapper = type('Wrapper', (), {})
    _module('gymnasium.spaces')
    _module('gymnasium.utils')

    pz = _module('pettingzoo')
    pz.AECEnv = _AECEnv
    pz.ParallelEnv = type('ParallelEnv', (), {})
    utils = _module('pettingzoo.utils')
    utils.agent_selector = _AgentSelector
    utils.BaseParallelWrapper = _PassThroughWrapper
    utils.BaseWrapper = _PassThroughWrapper
    pz.utils = utils
    wr = _module('pettingzoo.utils.wrappers')
    wr.OrderEnforcingWrapper = _PassThroughWrapper
    wr.BaseWrapper = _PassThroughWrapper
    wr.BaseParallelWrapper = _PassThroughWrapper
    utils.wrappers = wr
    conv = _module('pettingzoo.utils.conversions')
    conv.aec_to_parallel_wrapper = _AecToParallel
    utils.conversions = conv
    env = _module('pettingzoo.utils.env')
    env.ParallelEnv = type('ParallelEnv', (), {'__class_getitem__': classmethod(lambda cls, item: cls)})
    env.AECEnv = _AECEnv
    env.AgentID = env.ObsType = env.ActionType = object
    utils.env = env

    frr = _module('free_range_rust')
    frr.Space = _SpaceFactory()

    ss = _module('supersuit')
    _module('supersuit.utils')
    _module('supersuit.utils.base_aec_wrapper')
    _module('supersuit.utils.wrapper_chooser')
    ss.utils = sys.modules['supersuit.utils']

    sa = _module('sqlalchemy')
    for name in ('create_engine', 'Column', 'Integer', 'Text', 'Date', 'ForeignKey', 'Float', 'Boolean', 'String'):
        setattr(sa, name, lambda *a, **k: None)
    orm = _module('sqlalchemy.orm')
    orm.declarative_base = lambda *a, **k: type('Base', (), {})
    orm.relationship = lambda *a, **k: None
    orm.sessionmaker = lambda *a, **k: None
    sql = _module('sqlalchemy.sql')
    sql.func = types.SimpleNamespace(now=lambda: None)

    for name in ('pygame', 'imageio'):
        _module(name)


def raw(env):
    """Reach the reference raw_env under the pass-through wrappers."""
    inner = env
    while True:
        if hasattr(inner, 'aec_env'):
            inner = inner.aec_env
        elif isinstance(inner, _PassThroughWrapper):
            inner = object.__getattribute__(inner, 'env')
        else:
            return inner
