"""ctypes binding of the C-ABI in include/dbw_render.h (the drop-in boundary).  There is NO CPU fallback: if the
shared library is missing or a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DBW_RENDER_LIB: explicit path of another build of the same library (kernel experiments); never a fallback
LIB_PATH = os.environ.get('DBW_RENDER_LIB') or os.path.join(_HERE, 'libdbw_render.so')

ABI_VERSION = 7


class DbwRenderSettings(ctypes.Structure):
    _fields_ = [
        ('n_views', ctypes.c_int32), ('height', ctypes.c_int32), ('width', ctypes.c_int32),
        ('faces_per_pixel', ctypes.c_int32), ('n_verts', ctypes.c_int32), ('n_faces', ctypes.c_int32),
        ('n_maps', ctypes.c_int32), ('alpha_view_stride', ctypes.c_int32),
        ('fx', ctypes.c_float), ('fy', ctypes.c_float), ('px', ctypes.c_float), ('py', ctypes.c_float),
        ('sigma', ctypes.c_float), ('blur_radius', ctypes.c_float), ('z_clip', ctypes.c_float),
        ('proj_eps', ctypes.c_float), ('background', ctypes.c_float * 3),
        ('clip_inside', ctypes.c_int32), ('perspective_correct', ctypes.c_int32),
        ('clip_barycentric', ctypes.c_int32), ('detach_bary', ctypes.c_int32), ('verts_are_ndc', ctypes.c_int32),
        ('n_map_floats', ctypes.c_int32), ('maps_are_texels4', ctypes.c_int32), ('save_fragment_state', ctypes.c_int32),
        ('alpha_group', ctypes.c_int32), ('n_static_faces', ctypes.c_int32), ('view_rows', ctypes.c_void_p),
    ]


class DbwSceneGeometry(ctypes.Structure):
    _fields_ = [('n_blocks', ctypes.c_int32), ('verts_per_block', ctypes.c_int32), ('n_ground_verts', ctypes.c_int32),
                ('reserved', ctypes.c_int32)] \
             + [(n, ctypes.c_void_p) for n in ('sq_eta', 'sq_omega', 'sq_eps', 'S', 'R_6d', 'T', 'ground_verts',
                                               'R_6d_ground', 'T_ground')] \
             + [('ratio_block_scene', ctypes.c_float), ('scale_min', ctypes.c_float), ('S_world', ctypes.c_float),
                ('R_world', ctypes.c_float * 9), ('T_world', ctypes.c_float * 3)]


class DbwLossEpilogue(ctypes.Structure):
    _fields_ = [('env_rgba', ctypes.c_void_p), ('target', ctypes.c_void_p), ('g_env', ctypes.c_void_p), ('rec', ctypes.c_void_p),
                ('loss_partials', ctypes.c_void_p), ('n_partials', ctypes.c_int32), ('inv_count', ctypes.c_float)]


class DbwTexJob(ctypes.Structure):
    _fields_ = [('textures', ctypes.c_void_p), ('atlas', ctypes.c_void_p), ('g_textures', ctypes.c_void_p),
                ('n_maps', ctypes.c_int32), ('txt_size', ctypes.c_int32), ('p_left', ctypes.c_int32), ('p_right', ctypes.c_int32),
                ('decimate', ctypes.c_int32), ('stage', ctypes.c_int32)]


class DbwMapDesc(ctypes.Structure):
    _fields_ = [('offset', ctypes.c_int32), ('height', ctypes.c_int32), ('width', ctypes.c_int32),
                ('reserved', ctypes.c_int32)]


EXPORTS = ['dbw_abi_version', 'dbw_debug_generic_kernel_only', 'dbw_sizeof_settings', 'dbw_last_error', 'dbw_workspace_bytes', 'dbw_render_forward', 'dbw_render_forward_ex', 'dbw_render_backward',
           'dbw_render_forward_loss', 'dbw_render_backward_scaled',
           'dbw_composite_mse', 'dbw_composite_mse_backward', 'dbw_render_forward_host', 'dbw_host_arena_release', 'dbw_launch_count',
           'dbw_timing_enable', 'dbw_timing_read', 'dbw_timing_reset', 'dbw_scene_geometry_forward',
           'dbw_scene_geometry_backward', 'dbw_scene_geometry_forward_env', 'dbw_scene_geometry_backward_parts',
           'dbw_opacity_forward', 'dbw_opacity_backward', 'dbw_texture_prep_forward', 'dbw_texture_prep_backward',
           'dbw_texture_prep_forward_multi', 'dbw_texture_prep_backward_multi',
           'dbw_comm_create', 'dbw_comm_buffer', 'dbw_comm_ipc_handle', 'dbw_comm_connect', 'dbw_comm_all_reduce', 'dbw_comm_error', 'dbw_comm_destroy']

_lib = None


class DbwError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DbwError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           f'or `make -C differentiable-blocksworld_b200/csrc` (there is no CPU fallback)')
        L = ctypes.CDLL(LIB_PATH)
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.dbw_abi_version.restype = ctypes.c_int
        L.dbw_sizeof_settings.restype = ctypes.c_size_t
        L.dbw_last_error.restype = ctypes.c_char_p
        L.dbw_launch_count.restype = ctypes.c_uint64
        L.dbw_workspace_bytes.argtypes = [ctypes.POINTER(DbwRenderSettings), ctypes.POINTER(sz), ctypes.POINTER(sz)]
        L.dbw_render_forward.argtypes = [ctypes.POINTER(DbwRenderSettings)] + [vp] * 12 + [sz, vp]
        L.dbw_render_forward_ex.argtypes = [ctypes.POINTER(DbwRenderSettings)] + [vp] * 12 + [sz, vp, vp, vp]
        L.dbw_render_backward.argtypes = [ctypes.POINTER(DbwRenderSettings)] + [vp] * 11 + [sz] + [vp] * 5 + [sz, vp]
        L.dbw_render_forward_loss.argtypes = [ctypes.POINTER(DbwRenderSettings)] + [vp] * 12 + [sz, ctypes.POINTER(DbwLossEpilogue), vp]
        L.dbw_render_backward_scaled.argtypes = [ctypes.POINTER(DbwRenderSettings)] + [vp] * 11 + [sz] + [vp] * 6 + [sz, vp]
        L.dbw_composite_mse.argtypes = [ctypes.c_int32] * 3 + [vp] * 3 + [ctypes.c_float] + [vp] * 5
        L.dbw_composite_mse_backward.argtypes = [ctypes.c_int32] * 3 + [vp] * 3 + [ctypes.c_float] + [vp] * 5
        L.dbw_render_forward_host.argtypes = [ctypes.POINTER(DbwRenderSettings)] + [vp] * 5 + [sz] + [vp] * 6
        L.dbw_host_arena_release.restype = None
        L.dbw_scene_geometry_forward.argtypes = [ctypes.POINTER(DbwSceneGeometry), vp, vp]
        L.dbw_scene_geometry_backward.argtypes = [ctypes.POINTER(DbwSceneGeometry)] + [vp] * 8
        L.dbw_scene_geometry_forward_env.argtypes = [ctypes.POINTER(DbwSceneGeometry), vp, ctypes.c_int32, vp, vp]
        L.dbw_scene_geometry_backward_parts.argtypes = [ctypes.POINTER(DbwSceneGeometry)] + [vp] * 9
        L.dbw_opacity_forward.argtypes = [vp, vp, ctypes.c_float, ctypes.c_float, vp, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp]
        L.dbw_opacity_backward.argtypes = [vp, vp, ctypes.c_float, ctypes.c_float, vp, vp, ctypes.c_int32, vp, vp]
        L.dbw_texture_prep_forward.argtypes = [vp] + [ctypes.c_int32] * 5 + [vp, vp]
        L.dbw_texture_prep_backward.argtypes = [vp] + [ctypes.c_int32] * 5 + [vp, vp, vp]
        L.dbw_comm_create.argtypes = [ctypes.c_int32, ctypes.c_int32, sz, ctypes.POINTER(vp)]
        L.dbw_comm_buffer.argtypes = [vp, ctypes.POINTER(vp)]
        L.dbw_comm_ipc_handle.argtypes = [vp, vp]
        L.dbw_comm_connect.argtypes = [vp, vp]
        L.dbw_comm_all_reduce.argtypes = [vp, vp, sz, vp]
        L.dbw_comm_error.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
        L.dbw_comm_destroy.argtypes = [vp]
        L.dbw_debug_generic_kernel_only.restype = None
        L.dbw_debug_generic_kernel_only.argtypes = [ctypes.c_int]
        L.dbw_texture_prep_forward_multi.argtypes = [ctypes.POINTER(DbwTexJob), ctypes.c_int32, vp]
        L.dbw_texture_prep_backward_multi.argtypes = [ctypes.POINTER(DbwTexJob), ctypes.c_int32, vp]
        L.dbw_timing_enable.restype = None
        L.dbw_timing_reset.restype = None
        L.dbw_timing_read.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
        if L.dbw_sizeof_settings() != ctypes.sizeof(DbwRenderSettings):
            raise DbwError('DbwRenderSettings layout mismatch between the library and its ctypes mirror')
        if L.dbw_abi_version() != ABI_VERSION:
            raise DbwError(f'ABI mismatch: library {L.dbw_abi_version()} != binding {ABI_VERSION}')
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise DbwError(f'{what} failed: {lib().dbw_last_error().decode()}')


def launch_count():
    return int(lib().dbw_launch_count())


def kernel_time_ms(kind, K=0):
    """(total ms, launches) of the raster kernel `kind` (0 forward, 1 backward) recorded since the last reset."""
    tot, n = ctypes.c_double(0), ctypes.c_int(0)
    check(lib().dbw_timing_read(kind, K, ctypes.byref(tot), ctypes.byref(n)), 'dbw_timing_read')
    return tot.value, n.value
