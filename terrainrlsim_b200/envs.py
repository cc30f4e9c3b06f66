"""The reference's Python front door, batched: simAdapter/terrainRLSim.py (ActionSpace, TerrainRLSimWrapper :3-163,
getEnvsList / getEnv :165-208) over the registry args/envs.json.

Same names and call sequence as the reference; every per-env quantity gains a leading [E] axis because the wrapped
adapter (sim_adapter.BatchedSimAdapter) drives E characters at once: `step(action [E][A])` returns
`(observation [E][F], reward [E], done [E] bool, None)`.

The registry and the arg / character / motion / terrain files it points to are the reference's own data tree, located like
the reference does through the TERRAINRL_PATH environment variable (or the `root` argument).
"""
import json
import os

import numpy as np

from .api import TrlError
from .sim_adapter import BatchedSimAdapter


class ActionSpace(object):
    """Wrapper for the action space of an env (terrainRLSim.py:3-24)."""

    def __init__(self, action_space):
        self._minimum = np.array(action_space[0])
        self._maximum = np.array(action_space[1])

    def getMinimum(self):
        return self._minimum

    def getMaximum(self):
        return self._maximum

    @property
    def low(self):
        return self._minimum

    @property
    def high(self):
        return self._maximum


class TerrainRLSimWrapper(object):
    """terrainRLSim.py:26-163 for a batch of envs."""

    def __init__(self, sim, render=False, config=None):
        self._sim = sim
        self._render = render
        self._done = None
        n_act, n_obs = sim.getActionSpaceSize(), sim.getObservationSpaceSize()
        self._action_space = ActionSpace([[-1] * n_act, [1] * n_act])
        self._observation_space = ActionSpace([[-1] * n_obs, [1] * n_obs])
        self._config = config or {}

    def render(self):
        pass  # rendering (OpenGL) is outside this library

    def updateAction(self, action):
        self._sim.updateAction(np.asarray(action, dtype=np.float64))
        self._sim.handleUpdatedAction()

    def update(self):
        self._sim.update()
        fallen = self._sim.agentHasFallen()
        self._done = fallen if self._done is None else (self._done | fallen)

    def getObservation(self):
        return np.asarray(self._sim.getState())

    def step(self, action):
        self.updateAction(np.array(action, dtype="float64"))
        if self._config.get("control_return") is True:
            i = 0
            while (not self._sim.needUpdatedAction()) and i < 50:
                self._sim.update()
                i += 1
        else:
            self._sim.update()
        ob = self.getObservation()
        reward = self.calcRewards()
        fallen = self._sim.agentHasFallen()
        self._done = fallen if self._done is None else (self._done | fallen)
        return ob, reward, self._done, None

    def calcRewards(self):
        return self._sim.calcReward()

    def reset(self):
        self._sim.initEpoch()
        self._done = np.zeros(self._sim.num_envs, dtype=bool)
        return self.getObservation()

    def initEpoch(self):
        self.reset()

    def finish(self):
        self._sim.finish()

    def getActionSpace(self):
        return self._action_space

    @property
    def action_space(self):
        return self._action_space

    @property
    def observation_space(self):
        return self._observation_space

    def endOfEpoch(self):
        return self._done

    def init(self):
        self._sim.init()

    def getEnv(self):
        return self._sim

    def setRandomSeed(self, seed):
        self.getEnv().setRandomSeed(seed)


def _root(root=None):
    root = root or os.environ.get("TERRAINRL_PATH")
    if not root:
        raise TrlError("set TERRAINRL_PATH (or pass root=) to the reference's data tree (args/, data/)")
    return root


def getEnvsList(root=None):
    """args/envs.json: {env name: {"config_file": ..., "time_limit": ..., ...}} (terrainRLSim.py:165-178)."""
    with open(os.path.join(_root(root), "args", "envs.json")) as f:
        return json.load(f)


def getEnv(env_name, render=False, num_envs=1, root=None, device=0, world=None):
    """terrainRLSim.py:180-208 with `num_envs` characters behind one wrapper. Returns None for an unknown name, like the
    reference."""
    root = _root(root)
    env_data = getEnvsList(root)
    if env_name not in env_data or not isinstance(env_data[env_name], dict):
        print("Env: ", env_name, " not found. Check that you have the correct env name.")
        return None
    config_file = env_data[env_name]["config_file"]
    sim = BatchedSimAdapter(["train", "-arg_file=", os.path.join(root, config_file), "-relative_file_path=", root + "/"],
                            num_envs, device=device, world=world)
    sim.init()
    return TerrainRLSimWrapper(sim, render=render, config=env_data[env_name])
