"""agphys — B200-native batched physics step behind Assistive Gym's env surface."""
__all__ = ['capi', 'scene', 'sim', 'feeding_batch', 'envs']
