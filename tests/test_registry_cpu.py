"""The drop-in registers the reference's gym ids (gym_futbol/__init__.py:3-28) with the same kwargs.  gym is not
installed here, so a test-local stand-in of ``gym.envs.registration`` records the calls; when /root/reference is
present the reference's own ``__init__`` is imported over the same stand-in and the two registries are compared."""
import importlib
import os
import sys
import types

import pytest


@pytest.fixture()
def fake_gym(monkeypatch):
    gym = types.ModuleType("gym")
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registry = {}

    def register(id, entry_point=None, kwargs=None, **extra):
        assert id not in registry, "registered twice: %s" % id
        registry[id] = {"entry_point": entry_point, "kwargs": dict(kwargs or {}), "extra": extra}

    registration.register, registration.registry = register, registry
    envs.registration = registration
    gym.envs = envs
    for name, mod in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", registration)):
        monkeypatch.setitem(sys.modules, name, mod)
    return registry


def test_register_envs_registers_the_reference_ids(fake_gym):
    import gym_futbol_b200           # registers on first import when gym is importable (it is: the stand-in)
    fake_gym.clear()
    gym_futbol_b200.register_envs()
    assert set(fake_gym) == {"Futbol-v0", "Futbol-v1", "Futbol2v2-v1", "Futbol5v5-v1"}
    assert fake_gym["Futbol-v0"] == {"entry_point": "gym_futbol_b200.envs:FutbolEnv", "kwargs": {}, "extra": {}}
    for id_, n in (("Futbol-v1", 10), ("Futbol2v2-v1", 2), ("Futbol5v5-v1", 5)):
        assert fake_gym[id_] == {"entry_point": "gym_futbol_b200.envs_v1:Futbol", "kwargs": {"number_of_player": n}, "extra": {}}
    # every entry point resolves to a class with the reference's constructor keywords
    import inspect
    for rec in fake_gym.values():
        mod, cls = rec["entry_point"].split(":")
        klass = getattr(importlib.import_module(mod), cls)
        params = inspect.signature(klass.__init__).parameters
        assert all(k in params for k in rec["kwargs"])
    from gym_futbol_b200.envs import FutbolEnv
    from gym_futbol_b200.envs_v1 import Futbol
    for k in ("length", "width", "goal_size", "game_time", "player_speed", "shoot_speed", "Debug", "pressure_range",
              "one_goal_end", "action_as_int", "only_reward_goal", "random_opp"):     # envs/futbol_env.py:134-138
        assert k in inspect.signature(FutbolEnv.__init__).parameters, k
    for k in ("width", "height", "total_time", "debug", "number_of_player"):          # envs_v1/futbol_env.py:63-65
        assert k in inspect.signature(Futbol.__init__).parameters, k


@pytest.mark.skipif(not os.path.isfile("/root/reference/gym_futbol/__init__.py"), reason="needs /root/reference")
def test_registry_matches_the_reference_package(fake_gym, monkeypatch):
    """Import the reference's own gym_futbol/__init__.py over the stand-in and compare id -> kwargs.  Its
    'Futbol-extrahard-v0' points at a class that does not exist (envs/futbol_extrahard_env.py is empty): not mirrored."""
    src = open("/root/reference/gym_futbol/__init__.py").read()
    ns = {}
    exec(compile(src, "/root/reference/gym_futbol/__init__.py", "exec"), ns)      # only `register(...)` calls
    ref = {k: dict(v) for k, v in fake_gym.items()}
    import gym_futbol_b200
    fake_gym.clear()
    gym_futbol_b200.register_envs()
    ref.pop("Futbol-extrahard-v0")
    assert set(ref) == set(fake_gym)
    for id_, rec in ref.items():
        assert rec["kwargs"] == fake_gym[id_]["kwargs"], id_
        assert rec["entry_point"].replace("gym_futbol.", "gym_futbol_b200.") == fake_gym[id_]["entry_point"], id_
