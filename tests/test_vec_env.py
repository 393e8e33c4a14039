"""`AssistiveVecEnv` (SURVEY.md §8(b) fused vector path): host path on the CPU harness, device-tensor path on the GPU."""
import numpy as np
import pytest

from assistive_gym_b200.vec_env import AssistiveRLlibVectorEnv, AssistiveVecEnv


def _check_host_path(lib, env_id, obs_dim, **kw):
    v = AssistiveVecEnv(env_id, n_envs=3, seed=11, _lib=lib, **kw)
    obs = v.reset()
    assert obs.shape == (3, obs_dim)
    rng = np.random.default_rng(0)
    for _ in range(3):
        o, r, d, info = v.step(rng.uniform(-1, 1, size=(3, 7)).astype(np.float32))
        assert o.shape == (3, obs_dim) and r.shape == (3,) and d.dtype == bool and not d.any()
        assert np.all(np.isfinite(o)) and np.all(np.isfinite(r))
    # episode end: all envs finish together after 200 steps and the batch resets inside step()
    v._t = 199
    o, r, d, info = v.step(np.zeros((3, 7), dtype=np.float32))
    assert 'terminal_observation' in info and v._t == 0 and o.shape == (3, obs_dim)
    v.close()


def test_vec_env_host_path_feeding(emu_lib):
    _check_host_path(emu_lib, 'assistive_gym:FeedingJaco-v1', 25)


def test_vec_env_host_path_bed_bathing(emu_lib):
    _check_host_path(emu_lib, 'assistive_gym:BedBathingSawyer-v1', 24)


def test_vec_env_host_path_dressing(emu_lib):
    _check_host_path(emu_lib, 'assistive_gym:DressingPR2-v1', 24, toc_attempts=6)


def test_vec_env_double_buffered_reset(emu_lib):
    """The standby copy is re-randomised in the background and swapped in at the end of the episode."""
    v = AssistiveVecEnv('assistive_gym:FeedingJaco-v1', n_envs=3, seed=11, _lib=emu_lib, double_buffer=True)
    o0 = v.reset()
    first = v.env
    rng = np.random.default_rng(0)
    for _ in range(2):
        o, r, d, info = v.step(rng.uniform(-1, 1, size=(3, 7)).astype(np.float32))
    v._t = 199
    o, r, d, info = v.step(np.zeros((3, 7), dtype=np.float32))
    assert v.env is not first and v._standby is first and 'terminal_observation' in info
    assert o.shape == (3, 25) and np.all(np.isfinite(o)) and not np.array_equal(o, o0)      # a different draw
    o, r, d, info = v.step(np.zeros((3, 7), dtype=np.float32))
    assert np.all(np.isfinite(o)) and np.all(np.isfinite(r))
    v.close()


def test_rllib_vector_env_interface(emu_lib):
    """vector_reset / reset_at / vector_step / get_sub_environments as RLlib's VectorEnv drives them (learn.py:41)."""
    v = AssistiveRLlibVectorEnv('assistive_gym:FeedingJaco-v1', n_envs=3, seed=5, _lib=emu_lib)
    obs = v.vector_reset()
    assert len(obs) == 3 and obs[0].shape == (25,) and v.num_envs == 3 and v.get_sub_environments() == []
    o, r, d, infos = v.vector_step([v.action_space.sample() for _ in range(3)])
    assert len(o) == 3 and isinstance(r[0], float) and d == [False] * 3 and set(infos[0]) == {'total_force_on_human', 'task_success'}
    v.vec._t = 199
    o, r, d, infos = v.vector_step(np.zeros((3, 7)))
    assert d == [True] * 3
    first = v.reset_at(0)                   # the batch resets once ...
    assert v.vec._t == 0 and first.shape == (25,)
    again = [v.reset_at(i) for i in (1, 2)]  # ... and the other envs of the round read their rows of the same reset
    assert v.vec._t == 0 and all(a.shape == (25,) for a in again)
    assert np.array_equal(v.reset_at(0), v._obs[0]) and v.vec._t == 0       # a second reset of env 0 starts a new round
    v.vec.close()


@pytest.mark.gpu
def test_vec_env_device_tensors(gpu_lib):
    import torch
    v = AssistiveVecEnv('assistive_gym:FeedingJaco-v1', n_envs=64, seed=3)
    obs0 = v.reset()
    g = torch.Generator(device='cuda').manual_seed(0)
    host = AssistiveVecEnv('assistive_gym:FeedingJaco-v1', n_envs=64, seed=3)
    host.reset()
    for _ in range(3):
        a = torch.rand((64, 7), generator=g, device='cuda') * 2 - 1
        o, r, d, info = v.step(a)
        assert o.is_cuda and o.shape == (64, 25) and r.shape == (64,)
        oh, rh, dh, ih = host.step(a.cpu().numpy())
        torch.cuda.synchronize()
        assert np.array_equal(o.cpu().numpy(), oh) and np.array_equal(r.cpu().numpy(), rh)     # same kernels, same bits
    v.close(); host.close()
