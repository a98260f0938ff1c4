"""Synthetic workloads that need no asset files.

* ``kat_trace``  -- a small "known-answer" frame exercising every fixed-function feature of the hot path (clipping with
  VS re-execution, culling, points, lines of several widths, polygon modes, all depth functions, blending, every wrap
  and filter mode, mip-mapped and cube textures, MSAA).  Used for golden fixtures and GPU parity tests.
* ``soup_trace`` -- BASELINE config 4: N random textured triangles ("triangle soup").
"""
import math
import struct
import numpy as np

from . import trace as T
from .viewer import Camera, pack_uniforms_model, pack_uniforms_scene, pack_uniforms_material


class PCG32:
    """Small deterministic generator (PCG-XSH-RR) so that traces are identical across numpy versions."""

    def __init__(self, seed=0x5EED):
        self.state = 0
        self.inc = (54 << 1) | 1
        self._step()
        self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF
        self._step()

    def _step(self):
        self.state = (self.state * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF

    def u32(self):
        old = self.state
        self._step()
        x = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        r = old >> 59
        return ((x >> r) | (x << ((-r) & 31))) & 0xFFFFFFFF

    def f(self):
        return self.u32() / 4294967296.0

    def uniform(self, a, b):
        return a + (b - a) * self.f()


def value_noise_texture(size, seed, cells=8):
    """Seeded value-noise RGBA8 image (size x size), tileable."""
    rng = np.random.RandomState(seed)
    g = rng.rand(cells, cells, 4).astype(np.float32)
    u = (np.arange(size, dtype=np.float32) + 0.5) / size * cells
    i0 = np.floor(u).astype(int) % cells
    i1 = (i0 + 1) % cells
    f = (u - np.floor(u)).astype(np.float32)
    f = f * f * (3 - 2 * f)
    a = g[i0][:, i0] * (1 - f)[None, :, None] + g[i0][:, i1] * f[None, :, None]
    b = g[i1][:, i0] * (1 - f)[None, :, None] + g[i1][:, i1] * f[None, :, None]
    img = a * (1 - f)[:, None, None] + b * f[:, None, None]
    fine = rng.rand(size, size, 4).astype(np.float32) * 0.25
    out = np.clip((img * 0.75 + fine) * 255.0, 0, 255).astype(np.uint8)
    out[..., 3] = np.clip(96 + out[..., 3] // 2, 0, 255)
    return np.ascontiguousarray(out)


def _verts(pos, uv=None, normal=None):
    n = len(pos)
    v = np.zeros((n, 16), np.float32)
    v[:, 0:3] = pos
    if uv is not None:
        v[:, 4:6] = uv
    v[:, 8:11] = normal if normal is not None else np.array([0, 0, 1], np.float32)
    v[:, 12:15] = np.array([1, 0, 0], np.float32)
    return v


def _random_tris(rng, n, zmin, zmax, smin, smax, spread=1.4):
    """n triangles in view space in front of a camera looking down -z; some poke through the frustum planes."""
    pos = np.zeros((n * 3, 3), np.float32)
    uv = np.zeros((n * 3, 2), np.float32)
    for i in range(n):
        z = -rng.uniform(zmin, zmax)
        half = -z * math.tan(math.radians(30.0)) * spread
        c = np.array([rng.uniform(-half, half) * 1.3, rng.uniform(-half, half), z])
        s = rng.uniform(smin, smax) * (-z) * 0.2
        for k in range(3):
            pos[3 * i + k] = c + np.array([rng.uniform(-s, s), rng.uniform(-s, s), rng.uniform(-s, s) * 0.7])
            uv[3 * i + k] = (rng.uniform(-1.5, 2.5), rng.uniform(-1.5, 2.5))
    return pos, uv


def kat_trace(width=192, height=128, msaa=False, reverse_z=False, seed=7, mipmaps=True, tex_size=64):
    w = T.TraceWriter()
    rng = PCG32(seed)
    cam = Camera(60.0, float(width) / float(height), 0.5)
    cam.reverse_z = reverse_z
    cam.look_at((0, 0, 0), (0, 0, -1), (0, 1, 0))
    proj, view = cam.projection(), cam.view()
    eye4 = np.eye(4, dtype=np.float32)
    mvp = proj @ view
    light = (0.5, 1.0, 0.5)
    model_bytes = pack_uniforms_model(reverse_z, eye4, mvp, np.eye(3), eye4)
    scene_bytes = pack_uniforms_scene((0.4, 0.4, 0.4), (0, 0, 0), light, (0.8, 0.8, 0.8))

    color = w.create_texture(width, height, T.TextureType_2D, T.TextureFormat_RGBA8,
                             T.TextureUsage_AttachmentColor | T.TextureUsage_RendererOutput, False, msaa)
    w.tex_set_sampler(color, T.Filter_LINEAR, T.Filter_LINEAR)
    w.tex_init(color)
    depth = w.create_texture(width, height, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_AttachmentDepth,
                             False, msaa)
    w.tex_init(depth)
    fbo = w.create_fbo(False)
    w.fbo_color(fbo, color, 0)
    w.fbo_depth(fbo, depth)

    # textures: every wrap mode, nearest/linear/trilinear, one NPOT
    texs = []
    specs = [(T.Wrap_REPEAT, T.Filter_LINEAR, tex_size, False), (T.Wrap_MIRRORED_REPEAT, T.Filter_NEAREST, tex_size, False),
             (T.Wrap_CLAMP_TO_EDGE, T.Filter_LINEAR_MIPMAP_LINEAR if mipmaps else T.Filter_LINEAR, tex_size, mipmaps),
             (T.Wrap_CLAMP_TO_BORDER, T.Filter_LINEAR, tex_size, False),
             (T.Wrap_REPEAT, T.Filter_LINEAR_MIPMAP_NEAREST if mipmaps else T.Filter_LINEAR, tex_size * 2, mipmaps),
             (T.Wrap_REPEAT, T.Filter_NEAREST_MIPMAP_LINEAR if mipmaps else T.Filter_NEAREST, 48, mipmaps)]
    for i, (wrap, filt, size, mip) in enumerate(specs):
        t = w.create_texture(size, size, T.TextureType_2D, T.TextureFormat_RGBA8,
                             T.TextureUsage_Sampler | T.TextureUsage_UploadData, mip, False)
        w.tex_set_sampler(t, filt, T.Filter_LINEAR, wrap, wrap, T.Wrap_CLAMP_TO_EDGE,
                          T.Border_WHITE if i % 2 else T.Border_BLACK)
        w.tex_set_data(t, [value_noise_texture(size, 100 + i)])
        texs.append(t)
    shadow_ph = w.create_texture(1, 1, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_Sampler, False, False)
    w.tex_set_sampler(shadow_ph, T.Filter_NEAREST, T.Filter_NEAREST, T.Wrap_CLAMP_TO_BORDER, T.Wrap_CLAMP_TO_BORDER)
    w.tex_set_data(shadow_ph, [np.ones((1, 1), np.float32)])
    cube = w.create_texture(16, 16, T.TextureType_CUBE, T.TextureFormat_RGBA8,
                            T.TextureUsage_Sampler | T.TextureUsage_UploadData, False, False)
    w.tex_set_sampler(cube, T.Filter_LINEAR, T.Filter_LINEAR)
    w.tex_set_data(cube, [value_noise_texture(16, 200 + f, 4) for f in range(6)])

    b_model = w.create_block("UniformsModel", 256)
    b_scene = w.create_block("UniformsScene", 64)
    b_mat = w.create_block("UniformsMaterial", 48)
    w.block_data(b_model, model_bytes)
    w.block_data(b_scene, scene_bytes)
    MB, SB, TB = T.UniformBlock_Model, T.UniformBlock_Scene, T.UniformBlock_Material

    prog_basic = w.create_program(T.Shading_BaseColor, [])
    prog_bp_tex = w.create_program(T.Shading_BlinnPhong, ["ALBEDO_MAP"])
    prog_bp = w.create_program(T.Shading_BlinnPhong, [])
    prog_sky = w.create_program(T.Shading_Skybox, ["CUBE_MAP"])
    s_albedo = w.create_sampler("u_albedoMap", T.TextureType_2D, T.TextureFormat_RGBA8)
    s_shadow = w.create_sampler("u_shadowMap", T.TextureType_2D, T.TextureFormat_FLOAT32)
    s_cube = w.create_sampler("u_cubeMap", T.TextureType_CUBE, T.TextureFormat_RGBA8)
    w.sampler_tex(s_shadow, shadow_ph)
    w.sampler_tex(s_cube, cube)

    def states(**kw):
        rs = T.RenderStates()
        rs.depthTest = True
        rs.depthFunc = T.DepthFunc_GREATER if reverse_z else T.DepthFunc_LESS
        for k, v in kw.items():
            setattr(rs, k, v)
        return w.create_pipeline(rs)

    def material(base, light_on=True, point_size=1.0, spec=1.0):
        w.block_data(b_mat, pack_uniforms_material(light_on, False, False, point_size, spec, base))

    w.frame_begin()
    w.begin_pass(fbo, True, True, (0.1, 0.2, 0.3, 1.0), 0.0 if reverse_z else 1.0)
    w.viewport(0, 0, width, height)

    # 1. textured, depth-tested, back-face-culled triangles -- one draw per texture/wrap/filter
    for i, t in enumerate(texs):
        pos, uv = _random_tris(rng, 10, 0.8, 6.0, 0.5, 2.5)
        vao = w.create_vao(_verts(pos, uv), np.arange(len(pos), dtype=np.int32))
        w.sampler_tex(s_albedo, t)
        material((1, 1, 1, 1), light_on=(i % 2 == 0))
        w.draw(vao, prog_bp_tex, states(cullFace=(i % 3 == 0)), {MB: b_model, SB: b_scene, TB: b_mat},
               {T.MaterialTexType_ALBEDO: s_albedo, T.MaterialTexType_SHADOWMAP: s_shadow})
    # 2. huge triangles crossing several frustum planes (clipper, fan append order), untextured
    big = np.array([(-30, -4, -3), (30, -3, -3.5), (0, 25, -0.2), (-8, -20, -2), (9, 10, -0.1), (-9, 12, -40),
                    (0.2, 0.1, 0.4), (3, 0.3, -5), (-3, 2, -6)], np.float32)
    vao = w.create_vao(_verts(big), np.arange(9, dtype=np.int32))
    material((0.9, 0.6, 0.2, 1.0))
    w.draw(vao, prog_bp, states(), {MB: b_model, SB: b_scene, TB: b_mat}, {T.MaterialTexType_SHADOWMAP: s_shadow})
    # 3. every depth function on overlapping quads, depth writes off for half of them
    for fn in range(8):
        x0 = -2.2 + 0.55 * fn
        q = np.array([(x0, -1.2, -3.0 - 0.01 * fn), (x0 + 0.5, -1.2, -3.0), (x0 + 0.5, 1.2, -2.6), (x0, 1.2, -3.4)], np.float32)
        vao = w.create_vao(_verts(q), np.array([0, 1, 2, 0, 2, 3], np.int32))
        material((0.1 * fn, 1.0 - 0.1 * fn, 0.5, 1.0))
        w.draw(vao, prog_basic, states(depthFunc=fn, depthMask=(fn % 2 == 0)), {MB: b_model, TB: b_mat}, {})
    # 4. lines (widths 1, 2, 3.5) incl. clipped ones, and points of several sizes
    for lw in (1.0, 2.0, 3.5):
        pts = []
        for _ in range(12):
            z0, z1 = -rng.uniform(0.3, 5), -rng.uniform(0.3, 5)
            pts += [(rng.uniform(-4, 4), rng.uniform(-3, 3), z0), (rng.uniform(-4, 4), rng.uniform(-3, 3), z1)]
        vao = w.create_vao(_verts(np.array(pts, np.float32)), np.arange(len(pts), dtype=np.int32))
        material((1.0, 1.0, 0.2 * lw, 1.0))
        w.draw(vao, prog_basic, states(primitiveType=T.Primitive_LINE, lineWidth=lw), {MB: b_model, TB: b_mat}, {})
    for ps in (1.0, 4.0, 7.5):
        pts = [(rng.uniform(-3, 3), rng.uniform(-2, 2), -rng.uniform(0.6, 5)) for _ in range(10)]
        vao = w.create_vao(_verts(np.array(pts, np.float32)), np.arange(len(pts), dtype=np.int32))
        material((0.2, 1.0, 1.0, 1.0), point_size=ps)
        w.draw(vao, prog_basic, states(primitiveType=T.Primitive_POINT), {MB: b_model, TB: b_mat}, {})
    # 5. polygon modes LINE and POINT (wireframe path), culling on
    pos, uv = _random_tris(rng, 8, 1.0, 4.0, 1.0, 2.5)
    vao = w.create_vao(_verts(pos, uv), np.arange(len(pos), dtype=np.int32))
    material((1.0, 0.3, 0.9, 1.0))
    w.draw(vao, prog_basic, states(polygonMode=T.PolygonMode_LINE), {MB: b_model, TB: b_mat}, {})
    material((0.3, 1.0, 0.3, 1.0), point_size=1.0)
    w.draw(vao, prog_basic, states(polygonMode=T.PolygonMode_POINT, depthTest=False), {MB: b_model, TB: b_mat}, {})
    # 6. cube-mapped background cube without depth writes (skybox path)
    from .viewer import cube_mesh
    cv, ci = cube_mesh()
    vao = w.create_vao(cv, ci)
    rot = np.eye(4, dtype=np.float32)
    w.block_data(b_model, pack_uniforms_model(reverse_z, eye4, proj @ rot, np.eye(3), eye4))
    w.draw(vao, prog_sky, states(depthFunc=T.DepthFunc_GEQUAL if reverse_z else T.DepthFunc_LEQUAL, depthMask=False),
           {MB: b_model}, {T.MaterialTexType_CUBE: s_cube})
    w.block_data(b_model, model_bytes)
    # 7. alpha-blended overlapping triangles in submission order, several blend equations
    blends = [(T.BlendFunc_ADD, T.BlendFactor_SRC_ALPHA, T.BlendFactor_ONE_MINUS_SRC_ALPHA),
              (T.BlendFunc_ADD, T.BlendFactor_ONE, T.BlendFactor_ONE),
              (T.BlendFunc_REVERSE_SUBTRACT, T.BlendFactor_SRC_ALPHA, T.BlendFactor_ONE),
              (T.BlendFunc_MAX, T.BlendFactor_DST_COLOR, T.BlendFactor_ONE_MINUS_DST_ALPHA),
              (T.BlendFunc_MIN, T.BlendFactor_SRC_COLOR, T.BlendFactor_DST_ALPHA)]
    for i, (fn, sf, df) in enumerate(blends):
        pos, uv = _random_tris(rng, 6, 0.8, 3.0, 1.0, 3.0)
        vao = w.create_vao(_verts(pos, uv), np.arange(len(pos), dtype=np.int32))
        rs = T.RenderStates()
        rs.depthTest = True
        rs.depthFunc = T.DepthFunc_GREATER if reverse_z else T.DepthFunc_LESS
        rs.depthMask = False
        rs.blend = True
        rs.blendFuncRgb = rs.blendFuncAlpha = fn
        rs.set_blend_factor(sf, df)
        w.sampler_tex(s_albedo, texs[i % len(texs)])
        material((1, 1, 1, 1), light_on=False)
        w.draw(vao, prog_bp_tex, w.create_pipeline(rs), {MB: b_model, SB: b_scene, TB: b_mat},
               {T.MaterialTexType_ALBEDO: s_albedo, T.MaterialTexType_SHADOWMAP: s_shadow})
    w.end_pass()
    w.frame_end()
    w.wait_idle()
    w.readback(color, "color")
    w.readback(depth, "depth")
    return w


def soup_trace(n_tris=100000, width=1920, height=1080, seed=0x5EED, n_textures=8, tex_size=1024, mipmaps=True,
               msaa=False):
    """C4: random textured triangles, positions uniform in the frustum slab z in [1,50], log-uniform edge length
    0.5-64 px, UVs in [0,4) (REPEAT), triangle i uses texture i mod n_textures, Blinn-Phong ALBEDO_MAP, depth LESS."""
    w = T.TraceWriter()
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    cam = Camera(60.0, float(width) / float(height), 0.5)
    cam.look_at((0, 0, 0), (0, 0, -1), (0, 1, 0))
    proj = cam.projection()
    eye4 = np.eye(4, dtype=np.float32)
    color = w.create_texture(width, height, T.TextureType_2D, T.TextureFormat_RGBA8,
                             T.TextureUsage_AttachmentColor | T.TextureUsage_RendererOutput, False, msaa)
    w.tex_init(color)
    depth = w.create_texture(width, height, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_AttachmentDepth,
                             False, msaa)
    w.tex_init(depth)
    fbo = w.create_fbo(False)
    w.fbo_color(fbo, color, 0)
    w.fbo_depth(fbo, depth)
    texs = []
    for i in range(n_textures):
        t = w.create_texture(tex_size, tex_size, T.TextureType_2D, T.TextureFormat_RGBA8,
                             T.TextureUsage_Sampler | T.TextureUsage_UploadData, mipmaps, False)
        w.tex_set_sampler(t, T.Filter_LINEAR_MIPMAP_LINEAR if mipmaps else T.Filter_LINEAR, T.Filter_LINEAR,
                          T.Wrap_REPEAT, T.Wrap_REPEAT)
        w.tex_set_data(t, [value_noise_texture(tex_size, 300 + i, 16)])
        texs.append(t)
    shadow_ph = w.create_texture(1, 1, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_Sampler, False, False)
    w.tex_set_data(shadow_ph, [np.ones((1, 1), np.float32)])
    z = -rs.uniform(1.0, 50.0, n_tris).astype(np.float32)
    th = math.tan(math.radians(30.0))
    cx = rs.uniform(-1, 1, n_tris).astype(np.float32) * (-z) * th * (float(width) / height)
    cy = rs.uniform(-1, 1, n_tris).astype(np.float32) * (-z) * th
    edge_px = np.exp(rs.uniform(math.log(0.5), math.log(64.0), n_tris)).astype(np.float32)
    edge = edge_px * (-z) * th * 2.0 / height
    pos = np.zeros((n_tris, 3, 3), np.float32)
    ang = rs.uniform(0, 2 * math.pi, n_tris).astype(np.float32)
    for k in range(3):
        a = ang + k * (2 * math.pi / 3) + rs.uniform(-0.4, 0.4, n_tris).astype(np.float32)
        pos[:, k, 0] = cx + np.cos(a) * edge * 0.6
        pos[:, k, 1] = cy + np.sin(a) * edge * 0.6
        pos[:, k, 2] = z + rs.uniform(-0.5, 0.5, n_tris).astype(np.float32) * edge
    flip = rs.rand(n_tris) < 0.5                      # random winding
    pos[flip] = pos[flip][:, ::-1, :]
    uv = rs.uniform(0, 4, (n_tris, 3, 2)).astype(np.float32)
    b_model = w.create_block("UniformsModel", 256)
    b_scene = w.create_block("UniformsScene", 64)
    b_mat = w.create_block("UniformsMaterial", 48)
    w.block_data(b_model, pack_uniforms_model(False, eye4, proj, np.eye(3), eye4))
    w.block_data(b_scene, pack_uniforms_scene((0.5, 0.5, 0.5), (0, 0, 0), (0, 5, 0), (0.5, 0.5, 0.5)))
    w.block_data(b_mat, pack_uniforms_material(True, False, False, 1.0, 1.0, (1, 1, 1, 1)))
    prog = w.create_program(T.Shading_BlinnPhong, ["ALBEDO_MAP"])
    st = T.RenderStates()
    st.depthTest = True
    pipe = w.create_pipeline(st)
    s_shadow = w.create_sampler("u_shadowMap", T.TextureType_2D, T.TextureFormat_FLOAT32)
    w.sampler_tex(s_shadow, shadow_ph)
    draws = []
    for i in range(n_textures):                        # triangle i uses texture i mod n_textures: one draw per texture
        sel = np.arange(i, n_tris, n_textures)
        if len(sel) == 0:
            continue
        p = pos[sel].reshape(-1, 3)
        v = _verts(p, uv[sel].reshape(-1, 2))
        vao = w.create_vao(v, np.arange(len(p), dtype=np.int32))
        s = w.create_sampler("u_albedoMap", T.TextureType_2D, T.TextureFormat_RGBA8)
        w.sampler_tex(s, texs[i])
        draws.append((vao, s))
    w.frame_begin()
    w.begin_pass(fbo, True, True, (0, 0, 0, 1), 1.0)
    w.viewport(0, 0, width, height)
    MB, SB, TB = T.UniformBlock_Model, T.UniformBlock_Scene, T.UniformBlock_Material
    for vao, s in draws:
        w.draw(vao, prog, pipe, {MB: b_model, SB: b_scene, TB: b_mat},
               {T.MaterialTexType_ALBEDO: s, T.MaterialTexType_SHADOWMAP: s_shadow})
    w.end_pass()
    w.frame_end()
    w.wait_idle()
    w.readback(color, "color")
    w.readback(depth, "depth")
    return w


def stress_trace(kind, width=512, height=384, n_tris=600, msaa=False, seed=5):
    """Capacity stress cases for the binning / clipping arenas (the reference never drops geometry):
    kind "mid"     -- many mid-size triangles whose pixel range spans 20-60 16x16 tiles each (bin pressure)
    kind "clipped" -- every triangle pokes through the near plane and usually a side plane (clip-vertex / fan pressure)
    Flat-shaded (ShaderBasic), depth-tested, 8 draws with different colours."""
    w = T.TraceWriter()
    rs = np.random.RandomState(seed)
    cam = Camera(60.0, float(width) / float(height), 0.5)
    cam.look_at((0, 0, 0), (0, 0, -1), (0, 1, 0))
    mvp = cam.projection() @ cam.view()
    eye4 = np.eye(4, dtype=np.float32)
    color = w.create_texture(width, height, T.TextureType_2D, T.TextureFormat_RGBA8,
                             T.TextureUsage_AttachmentColor | T.TextureUsage_RendererOutput, False, msaa)
    w.tex_init(color)
    depth = w.create_texture(width, height, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_AttachmentDepth, False, msaa)
    w.tex_init(depth)
    fbo = w.create_fbo(False)
    w.fbo_color(fbo, color, 0)
    w.fbo_depth(fbo, depth)
    th = math.tan(math.radians(30.0))
    aspect = float(width) / height
    pos = np.zeros((n_tris, 3, 3), np.float32)
    if kind == "mid":
        z = -rs.uniform(2.0, 20.0, n_tris).astype(np.float32)
        cx = rs.uniform(-1, 1, n_tris).astype(np.float32) * (-z) * th * aspect
        cy = rs.uniform(-1, 1, n_tris).astype(np.float32) * (-z) * th
        edge_px = rs.uniform(70.0, 120.0, n_tris).astype(np.float32)          # 5..8 tiles per axis
        edge = edge_px * (-z) * th * 2.0 / height
        ang = rs.uniform(0, 2 * math.pi, n_tris).astype(np.float32)
        for k in range(3):
            a = ang + k * (2 * math.pi / 3)
            pos[:, k, 0] = cx + np.cos(a) * edge * 0.6
            pos[:, k, 1] = cy + np.sin(a) * edge * 0.6
            pos[:, k, 2] = z + rs.uniform(-0.2, 0.2, n_tris).astype(np.float32)
    elif kind == "clipped":
        for k in range(3):
            pos[:, k, 0] = rs.uniform(-6, 6, n_tris)
            pos[:, k, 1] = rs.uniform(-4, 4, n_tris)
        pos[:, 0, 2] = rs.uniform(0.1, 2.0, n_tris)       # behind the camera
        pos[:, 1, 2] = -rs.uniform(0.8, 4.0, n_tris)      # in front
        pos[:, 2, 2] = -rs.uniform(0.8, 4.0, n_tris)
    else:
        raise ValueError(kind)
    b_model = w.create_block("UniformsModel", 256)
    b_mat = w.create_block("UniformsMaterial", 48)
    w.block_data(b_model, pack_uniforms_model(False, eye4, mvp, np.eye(3), eye4))
    prog = w.create_program(T.Shading_BaseColor, [])
    st = T.RenderStates()
    st.depthTest = True
    pipe = w.create_pipeline(st)
    MB, TB = T.UniformBlock_Model, T.UniformBlock_Material
    n_draws = 8
    vaos = []
    for i in range(n_draws):
        p = pos[i::n_draws].reshape(-1, 3)
        vaos.append(w.create_vao(_verts(p), np.arange(len(p), dtype=np.int32)))
    w.frame_begin()
    w.begin_pass(fbo, True, True, (0.05, 0.05, 0.1, 1.0), 1.0)
    w.viewport(0, 0, width, height)
    for i, vao in enumerate(vaos):
        w.block_data(b_mat, pack_uniforms_material(False, False, False, 1.0, 1.0,
                                                   (0.2 + 0.1 * i, 1.0 - 0.11 * i, 0.3 + 0.05 * (i % 3), 1.0)))
        w.draw(vao, prog, pipe, {MB: b_model, TB: b_mat}, {})
    w.end_pass()
    w.frame_end()
    w.wait_idle()
    w.readback(color, "color")
    w.readback(depth, "depth")
    return w


def edge_trace(msaa=False):
    """Degenerate inputs the API must take in its stride: a render pass without draws (clear only), a draw with no
    indices, a 1x1 and a 3x2 framebuffer, a viewport smaller than / offset inside the framebuffer's tile grid, a pass
    that loads (no clear) what the previous pass wrote, degenerate (zero-area) and fully clipped triangles, a
    depth-only pass with zero primitives, and a blended draw as the FIRST draw of a pass."""
    w = T.TraceWriter()
    eye4 = np.eye(4, dtype=np.float32)
    MB, TB = T.UniformBlock_Model, T.UniformBlock_Material
    b_model = w.create_block("UniformsModel", 256)
    b_mat = w.create_block("UniformsMaterial", 48)
    w.block_data(b_model, pack_uniforms_model(False, eye4, eye4, np.eye(3), eye4))
    prog = w.create_program(T.Shading_BaseColor, [])
    rs = T.RenderStates()
    rs.depthTest = True
    rs.cullFace = False
    pipe = w.create_pipeline(rs)
    rb = T.RenderStates()
    rb.depthTest = False
    rb.cullFace = False
    rb.blend = True
    rb.set_blend_factor(T.BlendFactor_SRC_ALPHA, T.BlendFactor_ONE_MINUS_SRC_ALPHA)
    pipe_blend = w.create_pipeline(rb)
    tri = _verts(np.array([(-0.8, -0.7, 0.2), (0.9, -0.6, 0.4), (0.1, 0.8, 0.6)], np.float32))
    big = _verts(np.array([(-3, -3, 0.5), (3, -3, 0.5), (0, 3, 0.5)], np.float32))
    degenerate = _verts(np.array([(-0.5, -0.5, 0.3), (0.5, 0.5, 0.3), (0.0, 0.0, 0.3),        # zero area
                                  (5, 5, 0.5), (6, 5, 0.5), (5, 6, 0.5)], np.float32))          # outside the frustum
    vao_tri = w.create_vao(tri, np.arange(3, dtype=np.int32))
    vao_big = w.create_vao(big, np.arange(3, dtype=np.int32))
    vao_deg = w.create_vao(degenerate, np.arange(6, dtype=np.int32))
    vao_empty = w.create_vao(tri, np.zeros(0, np.int32))
    outs = []

    def target(wd, ht, tag):
        c = w.create_texture(wd, ht, T.TextureType_2D, T.TextureFormat_RGBA8,
                             T.TextureUsage_AttachmentColor | T.TextureUsage_RendererOutput, False, msaa)
        w.tex_init(c)
        d = w.create_texture(wd, ht, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_AttachmentDepth, False, msaa)
        w.tex_init(d)
        f = w.create_fbo(False)
        w.fbo_color(f, c, 0)
        w.fbo_depth(f, d)
        outs.append((c, "color_" + tag))
        outs.append((d, "depth_" + tag))
        return f

    def draw(vao, base, p=None):
        w.block_data(b_mat, pack_uniforms_material(False, False, False, 1.0, 1.0, base))
        w.draw(vao, prog, p if p is not None else pipe, {MB: b_model, TB: b_mat}, {})

    f_clear, f_1x1, f_3x2, f_vp, f_load = target(70, 50, "clear"), target(1, 1, "1x1"), target(3, 2, "3x2"), target(100, 60, "vp"), target(64, 64, "load")
    d_only = w.create_texture(40, 40, T.TextureType_2D, T.TextureFormat_FLOAT32, T.TextureUsage_AttachmentDepth, False, False)
    w.tex_init(d_only)
    f_depth = w.create_fbo(False)
    w.fbo_depth(f_depth, d_only)
    outs.append((d_only, "depth_only_empty"))

    w.frame_begin()
    w.begin_pass(f_clear, True, True, (0.2, 0.4, 0.6, 0.8), 1.0)          # no draws: clear only
    w.viewport(0, 0, 70, 50)
    w.end_pass()
    w.begin_pass(f_1x1, True, True, (0, 0, 0, 1), 1.0)
    w.viewport(0, 0, 1, 1)
    draw(vao_big, (1.0, 0.5, 0.25, 1.0))
    draw(vao_empty, (0, 1, 0, 1))                                         # indexCount == 0
    w.end_pass()
    w.begin_pass(f_3x2, True, True, (0, 0, 0, 1), 1.0)
    w.viewport(0, 0, 3, 2)
    draw(vao_tri, (0.3, 0.6, 0.9, 1.0))
    draw(vao_deg, (1, 0, 0, 1))                                           # zero-area + fully clipped
    w.end_pass()
    w.begin_pass(f_vp, True, True, (0.05, 0.05, 0.05, 1), 1.0)           # viewport smaller than the framebuffer
    w.viewport(0, 0, 37, 23)
    draw(vao_tri, (0.9, 0.8, 0.1, 1.0))
    w.end_pass()
    w.begin_pass(f_load, True, True, (0.1, 0.1, 0.1, 1), 1.0)
    w.viewport(0, 0, 64, 64)
    draw(vao_tri, (0.2, 0.9, 0.4, 1.0))
    w.end_pass()
    w.begin_pass(f_load, False, False, (0, 0, 0, 0), 1.0)                 # loads colour + depth; blended draw comes FIRST
    w.viewport(0, 0, 64, 64)
    draw(vao_big, (1.0, 0.2, 0.2, 0.5), pipe_blend)
    draw(vao_tri, (0.1, 0.1, 0.9, 1.0))
    w.end_pass()
    w.begin_pass(f_depth, False, True, (0, 0, 0, 0), 1.0)                 # depth-only pass, nothing to rasterise
    w.viewport(0, 0, 40, 40)
    draw(vao_deg, (0, 0, 0, 1))
    w.end_pass()
    w.frame_end()
    w.wait_idle()
    for t, tag in outs:
        w.readback(t, tag)
    return w
