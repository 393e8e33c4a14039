from .agents.human import Human
from .agents.robot import Sawyer
from .bed_bathing import BedBathingEnv

robot_arm = 'left'
human_controllable_joint_indices = list(range(0, 10))      # human.right_arm_joints (bed_bathing_envs.py)


class BedBathingSawyerEnv(BedBathingEnv):
    """`assistive_gym:BedBathingSawyer-v1` (reference envs/bed_bathing_envs.py)."""

    def __init__(self, n_envs=1, device=0, seed=1001, config=None):
        super().__init__(robot=Sawyer(robot_arm), human=Human(human_controllable_joint_indices, controllable=False),
                         n_envs=n_envs, device=device, seed=seed, config=config)
