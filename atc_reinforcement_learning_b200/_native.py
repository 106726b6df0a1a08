"""ctypes binding of the C ABI declared in include/atc_b200.h (the shared library built from csrc/atc_kernels.cu).

There is NO fallback: if the library is missing or fails to load, importing the env raises.  `build_library()` is
what __graft_entry__.build() calls; it cross-compiles for sm_100a with nvcc (no GPU needed)."""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libatc_b200.so')
INCLUDE = os.path.join(ROOT, 'include')

ABI_VERSION = 5
MAX_AIRCRAFT = 8
TEXT_MAX = 40
OBS_DIM = 10

EXPORTS = ['atc_abi_version', 'atc_compact_grid_budget', 'atc_create', 'atc_destroy', 'atc_reset', 'atc_step', 'atc_rollout', 'atc_step_host',
           'atc_rollout_host', 'atc_query_mva', 'atc_query_corridor', 'atc_launch_count', 'atc_last_error',
           'atc_obs_stats_update', 'atc_obs_normalize', 'atc_render', 'atc_last_launch_info', 'atc_vecnorm_run',
           'atc_vecnorm_last_error', 'atc_vecnorm_scratch_doubles', 'atc_vecnorm_max_steps', 'atc_render_text', 'atc_vecnorm_plan']

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class AtcSectorDesc(C.Structure):
    _fields_ = [
        ('n_mva', C.c_int32), ('n_vertices', C.c_int32),
        ('ring_xy', _dp), ('ring_off', _ip), ('mva_height', _dp), ('mva_bounds', _dp),
        ('runway_x', C.c_double), ('runway_y', C.c_double), ('runway_h', C.c_double), ('phi_to_runway', C.c_double),
        ('faf', C.c_double * 2), ('normal', C.c_double * 2),
        ('tri_h', C.c_double * 8), ('tri_1', C.c_double * 8), ('tri_2', C.c_double * 8),
        ('sin_to_runway', C.c_double), ('cos_to_runway', C.c_double), ('glide_tan', C.c_double),
        ('bbox', C.c_double * 4), ('world_max_distance', C.c_double), ('faf_mva', C.c_double),
        ('norm_min', C.c_float * OBS_DIM), ('norm_max', C.c_float * OBS_DIM),
        ('n_entry', C.c_int32), ('entry_xyphi', _dp), ('level_off', _ip), ('levels', _ip),
        ('grid_nx', C.c_int32), ('grid_ny', C.c_int32), ('grid_inv_cell', C.c_double),
        ('grid_x0', C.c_double), ('grid_y0', C.c_double),
        ('grid_cell', C.POINTER(C.c_uint16)), ('n_mixed', C.c_int32), ('n_prog', C.c_int32),
        ('grid_prog_off', C.POINTER(C.c_uint32)), ('grid_prog', C.POINTER(C.c_uint16)), ('grid_line', _dp),
        ('wind_gx', C.c_int32), ('wind_gy', C.c_int32), ('wind', _fp),
        ('cgrid_nx', C.c_int32), ('cgrid_ny', C.c_int32), ('cgrid_inv_cell', C.c_double),
        ('cgrid_x0', C.c_double), ('cgrid_y0', C.c_double), ('cgrid_cell', C.POINTER(C.c_uint16)),
        ('cgrid_n_blocks', C.c_int32), ('n_cline', C.c_int32), ('cline', _dp),
    ]


class AtcSimParams(C.Structure):
    _fields_ = [
        ('timestep', C.c_double), ('reward_shaping', C.c_int32), ('normalize_state', C.c_int32),
        ('discrete_action_space', C.c_int32), ('normalize_reset_obs', C.c_int32), ('n_env', C.c_int32),
        ('n_aircraft', C.c_int32), ('track_actions', C.c_int32), ('exact_math', C.c_int32), ('seed', C.c_uint64),
        ('env_index_base', C.c_int64),
    ]


class AtcBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('state', 'last_action', 'timesteps', 'episodes', 'ep_return',
                                          'actions_taken', 'last_ep_return', 'last_ep_len', 'win_ring')]


class AtcStepIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('actions', 'obs', 'raw_obs', 'reward', 'done', 'term')]


class AtcVecNormState(C.Structure):
    _fields_ = [('obs_rms', C.c_void_p), ('ret_rms', C.c_void_p), ('ret', C.c_void_p), ('scratch', C.c_void_p),
                ('scratch_doubles', C.c_int64), ('sync', C.c_void_p), ('nonfinite', C.c_void_p)]


class AtcVecNormParams(C.Structure):
    _fields_ = [('training', C.c_int32), ('norm_obs', C.c_int32), ('norm_reward', C.c_int32), ('reserved', C.c_int32),
                ('clip_obs', C.c_double), ('clip_reward', C.c_double), ('gamma', C.c_double), ('epsilon', C.c_double)]


class AtcLaunchInfo(C.Structure):
    _fields_ = [('kernel', C.c_int32), ('n_steps', C.c_int32), ('grid', C.c_int32), ('block', C.c_int32),
                ('pairs_per_cta', C.c_int32), ('lanes_per_env', C.c_int32), ('wind', C.c_int32),
                ('track_actions', C.c_int32), ('exact_math', C.c_int32), ('raw_obs', C.c_int32),
                ('cfg', C.c_int32), ('reserved', C.c_int32), ('dyn_smem_bytes', C.c_int64)]


KERNEL_NAMES = {0: 'none', 1: 'atc_step_kernel', 2: 'atc_rollout_pipe_kernel', 3: 'atc_rollout_pipe_kernel'}


SOURCES = ['atc_kernels.cu', 'atc_vecnorm.cu', 'atc_text.cu']


def nvcc_command(out=LIB_PATH):
    return ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
            '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static', '-I', INCLUDE, '-o', out] + \
           [os.path.join(CSRC, f) for f in SOURCES]


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libatc_b200.so (in-tree, travels to the GPU box)."""
    src = [os.path.join(CSRC, f) for f in SOURCES] + [os.path.join(INCLUDE, 'atc_b200.h')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src):
        return LIB_PATH
    if shutil.which('nvcc') is None:
        raise RuntimeError('nvcc not found: cannot build %s' % LIB_PATH)
    cmd = nvcc_command()
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises (never falls back) when it is not built or cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('%s is missing — run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(there is no CPU fallback)' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RuntimeError('%s does not export %s' % (LIB_PATH, name))
    L.atc_abi_version.restype = C.c_int
    if L.atc_abi_version() != ABI_VERSION:
        raise RuntimeError('ABI mismatch: library %d, binding %d' % (L.atc_abi_version(), ABI_VERSION))
    vp = C.c_void_p
    L.atc_create.argtypes = [C.POINTER(AtcSectorDesc), C.POINTER(AtcSimParams), C.c_int, C.POINTER(vp)]
    L.atc_destroy.argtypes = [vp]
    L.atc_reset.argtypes = [vp, C.POINTER(AtcBuffers), vp, vp, vp, vp]
    L.atc_step.argtypes = [vp, C.POINTER(AtcBuffers), C.POINTER(AtcStepIO), C.c_int, vp]
    L.atc_rollout.argtypes = [vp, C.POINTER(AtcBuffers), C.POINTER(AtcStepIO), C.c_int, vp]
    L.atc_step_host.argtypes = [vp, C.POINTER(AtcBuffers), C.POINTER(AtcStepIO), C.POINTER(AtcStepIO), C.c_int, vp]
    L.atc_rollout_host.argtypes = [vp, C.POINTER(AtcBuffers), C.POINTER(AtcStepIO), C.POINTER(AtcStepIO), C.c_int, vp]
    L.atc_query_mva.argtypes = [vp, C.c_int, vp, vp, vp]
    L.atc_query_corridor.argtypes = [vp, C.c_int, vp, vp, vp]
    L.atc_obs_stats_update.argtypes = [vp, C.c_int64, C.c_int32, vp, vp, vp, vp]
    L.atc_obs_normalize.argtypes = [vp, C.c_int64, C.c_int32, vp, C.c_double, C.c_double, vp, vp]
    L.atc_obs_stats_update.restype = C.c_int
    L.atc_obs_normalize.restype = C.c_int
    L.atc_vecnorm_run.argtypes = [C.POINTER(AtcVecNormState), C.POINTER(AtcVecNormParams), C.c_int32, C.c_int64, C.c_int32,
                                  vp, vp, vp, vp, vp, C.c_int, vp]
    L.atc_vecnorm_run.restype = C.c_int
    L.atc_vecnorm_scratch_doubles.argtypes = [C.c_int, C.c_int32, C.c_int64, C.c_int32]
    L.atc_vecnorm_scratch_doubles.restype = C.c_int64
    L.atc_vecnorm_max_steps.argtypes = [C.c_int]
    L.atc_vecnorm_max_steps.restype = C.c_int32
    L.atc_vecnorm_plan.argtypes = [C.c_int, C.c_int, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]
    L.atc_vecnorm_plan.restype = C.c_int
    L.atc_vecnorm_last_error.argtypes = []
    L.atc_vecnorm_last_error.restype = C.c_char_p
    L.atc_render.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp]
    L.atc_render.restype = C.c_int
    L.atc_render_text.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.atc_render_text.restype = C.c_int
    L.atc_compact_grid_budget.argtypes = []
    L.atc_compact_grid_budget.restype = C.c_int64
    L.atc_launch_count.argtypes = [vp]
    L.atc_launch_count.restype = C.c_int64
    L.atc_last_launch_info.argtypes = [vp, C.POINTER(AtcLaunchInfo)]
    L.atc_last_launch_info.restype = C.c_int
    L.atc_last_error.argtypes = [vp]
    L.atc_last_error.restype = C.c_char_p
    for name in ('atc_create', 'atc_destroy', 'atc_reset', 'atc_step', 'atc_rollout', 'atc_step_host',
                 'atc_rollout_host', 'atc_query_mva', 'atc_query_corridor'):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _np_ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def sector_desc(cs):
    """CompiledSector -> AtcSectorDesc (host pointers into the CompiledSector's numpy arrays; keep `cs` alive)."""
    d = AtcSectorDesc()
    d.n_mva, d.n_vertices = len(cs.rings), len(cs.ring_xy)
    d.ring_xy, d.ring_off = _np_ptr(cs.ring_xy, C.c_double), _np_ptr(cs.ring_off, C.c_int32)
    d.mva_height, d.mva_bounds = _np_ptr(cs.mva_height, C.c_double), _np_ptr(cs.mva_bounds, C.c_double)
    d.runway_x, d.runway_y, d.runway_h = cs.runway
    d.phi_to_runway = cs.phi_to_runway
    d.faf[:] = cs.faf
    d.normal[:] = cs.normal
    d.tri_h[:] = cs.tri_h.ravel().tolist()
    d.tri_1[:] = cs.tri_1.ravel().tolist()
    d.tri_2[:] = cs.tri_2.ravel().tolist()
    d.sin_to_runway, d.cos_to_runway, d.glide_tan = cs.sin_to_runway, cs.cos_to_runway, cs.glide_tan
    d.bbox[:] = cs.bbox.tolist()
    d.world_max_distance, d.faf_mva = cs.world_max_distance, cs.faf_mva
    d.norm_min[:] = cs.norm_min.tolist()
    d.norm_max[:] = cs.norm_max.tolist()
    d.n_entry = len(cs.entry_xyphi)
    d.entry_xyphi, d.level_off = _np_ptr(cs.entry_xyphi, C.c_double), _np_ptr(cs.level_off, C.c_int32)
    d.levels = _np_ptr(cs.levels, C.c_int32)
    d.grid_nx, d.grid_ny, d.grid_inv_cell = cs.grid_nx, cs.grid_ny, cs.grid_inv_cell
    d.grid_x0, d.grid_y0 = cs.grid_x0, cs.grid_y0
    d.grid_cell = _np_ptr(cs.grid_cell, C.c_uint16)
    d.n_mixed, d.n_prog = len(cs.grid_prog_off), len(cs.grid_prog)
    d.grid_prog_off, d.grid_prog = _np_ptr(cs.grid_prog_off, C.c_uint32), _np_ptr(cs.grid_prog, C.c_uint16)
    d.grid_line = _np_ptr(cs.grid_line, C.c_double)
    if cs.wind is not None:
        d.wind_gy, d.wind_gx = cs.wind.shape[0], cs.wind.shape[1]
        d.wind = _np_ptr(cs.wind, C.c_float)
    cg = getattr(cs, 'compact', None)
    if cg is not None:
        d.cgrid_nx, d.cgrid_ny, d.cgrid_inv_cell = cg.grid_nx, cg.grid_ny, cg.sub_inv_cell
        d.cgrid_x0, d.cgrid_y0 = cg.grid_x0, cg.grid_y0
        d.cgrid_cell = _np_ptr(cg.grid_cell, C.c_uint16)
        d.cgrid_n_blocks = cg.n_blocks
        d.n_cline, d.cline = cg.n_lines, _np_ptr(cg.lines, C.c_double)
    return d


class AtcError(RuntimeError):
    pass


def check(handle, rc):
    if rc != 0:
        msg = lib().atc_last_error(handle)
        raise AtcError('atc_b200 error %d: %s' % (rc, msg.decode() if msg else '?'))
