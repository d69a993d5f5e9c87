"""Agent protocol driving a PSRS environment (offsim4rl/agents/agent.py:6-25)."""
import abc


class Agent(abc.ABC):
    @property
    @abc.abstractmethod
    def action_dist_type(self):
        ...

    @abc.abstractmethod
    def begin_episode(self, observation):
        """Returns the action distribution for the first observation."""

    @abc.abstractmethod
    def commit_action(self, action):
        """Told which action the simulator realised."""

    @abc.abstractmethod
    def step(self, prev_reward, observation):
        """Returns the action distribution for the next observation."""

    @abc.abstractmethod
    def end_episode(self, reward, truncated=False):
        ...
