from .agent import Agent


class Furniture(Agent):
    """Wheelchair, table, bowl, bed (reference envs/agents/furniture.py)."""
    pass
