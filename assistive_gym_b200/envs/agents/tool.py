from .agent import Agent


class Tool(Agent):
    """Spoon / cup / wiper held by the gripper (reference envs/agents/tool.py)."""
    pass
