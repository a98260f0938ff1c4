"""Headless "Viewer": emits the Renderer-API call sequence of one SoftGLRender frame.

This mirrors the *caller* of the hot path -- src/Viewer/Viewer.cpp (drawFrame :100-141,
drawShadowMap :143-168, drawScene :293-328, pipelineDraw :388-396, uniform updates
:673-716), src/Viewer/Environment.cpp (IBL passes :25-195), src/Viewer/QuadFilter.cpp
(FXAA pass :12-112) and src/Viewer/ModelLoader.cpp (floor/axis/light meshes :48-115) --
and records the calls into a trace instead of executing them, so the same frame can be
replayed on RendererSoft (oracle) and RendererCUDA.  It is host-side scene logic, not
part of the accelerated path.
"""
import math
import os
import struct
import numpy as np

from . import trace as T
from . import assets as A

SHADOW_MAP_SIZE = 512            # Viewer.cpp:16-17
IRRADIANCE_SIZE = 32             # Environment.h:20
PREFILTER_SIZE = 128             # Environment.h:22
PREFILTER_LEVELS = 5             # Environment.h:21
CAMERA_FOV, CAMERA_NEAR = 60.0, 0.01

AA_NONE, AA_MSAA, AA_FXAA = 0, 1, 2


class Config:
    """Twin of View::Config (src/Viewer/Config.h:25-55)."""

    def __init__(self, **kw):
        self.wireframe = False
        self.world_axis = True
        self.show_skybox = False
        self.show_floor = True
        self.shadow_map = True
        self.pbr_ibl = False
        self.mipmaps = False
        self.cull_face = True
        self.depth_test = True
        self.reverse_z = False
        self.clear_color = (0.0, 0.0, 0.0, 0.0)
        self.ambient_color = (0.5, 0.5, 0.5)
        self.show_light = True
        ang = math.radians(235.0)                              # ConfigPanel.h:68, ConfigPanel.cpp:218-220
        self.point_light_position = (2.0 * math.sin(ang), 2.0 * 1.2, 2.0 * math.cos(ang))
        self.point_light_color = (0.5, 0.5, 0.5)
        self.aa_type = AA_NONE
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _normalize(v):
    v = _f32(v)
    return v / np.float32(np.linalg.norm(v))


class Camera:
    """src/Viewer/Camera.cpp:26-76 (infinite-far perspective, optional reversed-Z)."""

    def __init__(self, fov_deg=CAMERA_FOV, aspect=1.0, near=CAMERA_NEAR):
        self.fov = math.radians(fov_deg)
        self.aspect = aspect
        self.near = near
        self.reverse_z = False
        self.eye = _f32([0, 0, 0])
        self.center = _f32([0, 0, -1])
        self.up = _f32([0, 1, 0])

    def look_at(self, eye, center, up):
        self.eye, self.center, self.up = _f32(eye), _f32(center), _f32(up)

    def projection(self):
        t = np.float32(1.0) / np.float32(math.tan(self.fov * 0.5))
        p = np.zeros((4, 4), np.float32)           # row-major maths view; p[r][c]
        p[0, 0] = t / np.float32(self.aspect)
        p[1, 1] = t
        if self.reverse_z:
            p[2, 2] = 0.0
            p[3, 2] = -1.0
            p[2, 3] = self.near
        else:
            p[2, 2] = -1.0
            p[3, 2] = -1.0
            p[2, 3] = -self.near
        return p

    def view(self):
        f = _normalize(self.center - self.eye)
        s = _normalize(np.cross(f, self.up))
        u = np.cross(s, f).astype(np.float32)
        v = np.eye(4, dtype=np.float32)
        v[0, :3] = s
        v[1, :3] = u
        v[2, :3] = -f
        v[0, 3] = -np.dot(s, self.eye)
        v[1, 3] = -np.dot(u, self.eye)
        v[2, 3] = np.dot(f, self.eye)
        return v


def _mat_bytes(m):
    """GLM column-major storage of a maths-convention (row,col) matrix."""
    return np.ascontiguousarray(_f32(m).T).tobytes()


def pack_uniforms_model(reverse_z, model, mvp, inv_t3, shadow_mvp):
    """UniformsModel (Material.h:70-76): u32 @0, mat4 @16, mat4 @80, mat3 (3 x aligned vec3) @144, mat4 @192."""
    b = bytearray(256)
    struct.pack_into("<I", b, 0, 1 if reverse_z else 0)
    b[16:80] = _mat_bytes(model)
    b[80:144] = _mat_bytes(mvp)
    it = np.zeros((3, 4), np.float32)
    it[:, :3] = _f32(inv_t3).T                     # three columns, 16-byte stride
    b[144:192] = it.tobytes()
    b[192:256] = _mat_bytes(shadow_mvp)
    return bytes(b)


def pack_uniforms_scene(ambient, cam_pos, light_pos, light_color):
    b = bytearray(64)
    for i, v in enumerate((ambient, cam_pos, light_pos, light_color)):
        struct.pack_into("<3f", b, 16 * i, *[float(x) for x in v])
    return bytes(b)


def pack_uniforms_material(enable_light, enable_ibl, enable_shadow, point_size, k_specular, base_color):
    b = bytearray(48)
    struct.pack_into("<3I2f", b, 0, int(enable_light), int(enable_ibl), int(enable_shadow),
                     float(point_size), float(k_specular))
    struct.pack_into("<4f", b, 32, *[float(x) for x in base_color])
    return bytes(b)


_SAMPLER_NAME = {                                  # Material.cpp:69-88
    T.MaterialTexType_ALBEDO: "u_albedoMap", T.MaterialTexType_NORMAL: "u_normalMap",
    T.MaterialTexType_EMISSIVE: "u_emissiveMap", T.MaterialTexType_AMBIENT_OCCLUSION: "u_aoMap",
    T.MaterialTexType_METAL_ROUGHNESS: "u_metalRoughnessMap", T.MaterialTexType_CUBE: "u_cubeMap",
    T.MaterialTexType_EQUIRECTANGULAR: "u_equirectangularMap",
    T.MaterialTexType_IBL_IRRADIANCE: "u_irradianceMap", T.MaterialTexType_IBL_PREFILTER: "u_prefilterMap",
    T.MaterialTexType_QUAD_FILTER: "u_screenTexture", T.MaterialTexType_SHADOWMAP: "u_shadowMap"}
_SAMPLER_DEFINE = {                                # Material.cpp:49-66
    T.MaterialTexType_ALBEDO: "ALBEDO_MAP", T.MaterialTexType_NORMAL: "NORMAL_MAP",
    T.MaterialTexType_EMISSIVE: "EMISSIVE_MAP", T.MaterialTexType_AMBIENT_OCCLUSION: "AO_MAP",
    T.MaterialTexType_METAL_ROUGHNESS: "METALROUGHNESS_MAP", T.MaterialTexType_CUBE: "CUBE_MAP",
    T.MaterialTexType_EQUIRECTANGULAR: "EQUIRECTANGULAR_MAP", T.MaterialTexType_IBL_IRRADIANCE: "IBL_MAP",
    T.MaterialTexType_IBL_PREFILTER: "IBL_MAP"}
_TEXKEY = {"albedo": T.MaterialTexType_ALBEDO, "normal": T.MaterialTexType_NORMAL,
           "emissive": T.MaterialTexType_EMISSIVE, "ao": T.MaterialTexType_AMBIENT_OCCLUSION,
           "metal_roughness": T.MaterialTexType_METAL_ROUGHNESS}


class Material:
    def __init__(self, shading, base_color=(1, 1, 1, 1), point_size=1.0, line_width=1.0,
                 double_sided=False, alpha_blend=False):
        self.shading = shading
        self.base_color = base_color
        self.point_size = point_size
        self.line_width = line_width
        self.double_sided = double_sided
        self.alpha_blend = alpha_blend
        self.texture_data = {}      # MaterialTexType -> dict(images=[...], wrap_u, wrap_v, wrap_w, cube)
        self.textures = {}          # MaterialTexType -> (tex id, type, format)
        self.defines = set()
        self.obj = None             # dict(shading, program, pipeline, blocks{}, samplers{})


class Drawable:
    def __init__(self, vertices, indices, primitive, material):
        self.vertices = vertices
        self.indices = indices
        self.primitive = primitive
        self.material = material
        self.vao = None


def cube_mesh():
    """12-triangle unit cube (src/Viewer/Cube.h:14-55 via ModelLoader::loadCubeMesh :29-45)."""
    f = [(-1, 1, -1), (-1, -1, -1), (1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1),
         (-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, -1), (-1, 1, 1), (-1, -1, 1),
         (1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, 1), (1, 1, -1), (1, -1, -1),
         (-1, -1, 1), (-1, 1, 1), (1, 1, 1), (1, 1, 1), (1, -1, 1), (-1, -1, 1),
         (-1, 1, -1), (1, 1, -1), (1, 1, 1), (1, 1, 1), (-1, 1, 1), (-1, 1, -1),
         (-1, -1, -1), (-1, -1, 1), (1, -1, -1), (1, -1, -1), (-1, -1, 1), (1, -1, 1)]
    pos = np.array(f, np.float32)
    return A.make_vertices(pos), np.arange(36, dtype=np.int32)


class Scene:
    """DemoScene (Model.h:108-124) + the procedural furniture of ModelLoader.cpp:48-115."""

    def __init__(self, config, model, skybox=None):
        self.model = model
        # world axis: 66 lines at y = -0.01
        pos = []
        for i in range(-16, 17):
            z = np.float32(0.2) * np.float32(i)
            pos += [(-3.2, -0.01, z), (3.2, -0.01, z), (z, -0.01, -3.2), (z, -0.01, 3.2)]
        self.world_axis = Drawable(A.make_vertices(np.array(pos, np.float32)),
                                   np.arange(len(pos), dtype=np.int32), T.Primitive_LINE,
                                   Material(T.Shading_BaseColor, (0.25, 0.25, 0.25, 1.0)))
        self.point_light = Drawable(A.make_vertices(np.array([config.point_light_position], np.float32)),
                                    np.array([0], np.int32), T.Primitive_POINT,
                                    Material(T.Shading_BaseColor, tuple(config.point_light_color) + (1.0,),
                                             point_size=10.0))
        fp = np.array([(-2, 0.01, 2), (-2, 0.01, -2), (2, 0.01, -2), (2, 0.01, 2)], np.float32)
        fuv = np.array([(0, 1), (0, 0), (1, 0), (1, 1)], np.float32)
        fn = np.array([(0, 1, 0)] * 4, np.float32)
        self.floor = Drawable(A.make_vertices(fp, fuv, fn), np.array([0, 2, 1, 0, 3, 2], np.int32),
                              T.Primitive_TRIANGLE, Material(T.Shading_BlinnPhong, double_sided=True))
        cv, ci = cube_mesh()
        self.skybox = Drawable(cv, ci, T.Primitive_TRIANGLE, Material(T.Shading_Skybox))
        self.skybox.ibl_ready = False
        if skybox is not None:
            self.skybox.material.texture_data = skybox
        # model meshes -> drawables with materials
        self.mesh_drawables = {}

    def drawable_for(self, mesh):
        d = self.mesh_drawables.get(id(mesh))
        if d is None:
            mat = Material(T.Shading_PBR if mesh.shading == "pbr" else T.Shading_BlinnPhong,
                           base_color=mesh.base_color, double_sided=mesh.double_sided,
                           alpha_blend=mesh.alpha_blend)
            for name, td in mesh.textures.items():
                mat.texture_data[_TEXKEY[name]] = dict(images=[td["image"]], wrap_u=td["wrap_u"],
                                                       wrap_v=td["wrap_v"], wrap_w=T.Wrap_REPEAT, cube=False)
            d = Drawable(mesh.vertices, mesh.indices, T.Primitive_TRIANGLE, mat)
            self.mesh_drawables[id(mesh)] = d
        return d


def load_skybox(path, image_cache=None):
    """ModelLoader::loadSkybox (ModelLoader.cpp:117-176): directory => 6-face cube, file => equirect."""
    cache = {} if image_cache is None else image_cache

    def img(p):
        if p not in cache:
            cache[p] = A.load_image_rgba(p)
        return cache[p]
    clamp = T.Wrap_CLAMP_TO_EDGE
    if path.endswith("/"):
        faces = [img(path + n + ".jpg") for n in ("right", "left", "top", "bottom", "front", "back")]
        return {T.MaterialTexType_CUBE: dict(images=faces, wrap_u=clamp, wrap_v=clamp, wrap_w=clamp, cube=True)}
    return {T.MaterialTexType_EQUIRECTANGULAR: dict(images=[img(path)], wrap_u=clamp, wrap_v=clamp,
                                                    wrap_w=clamp, cube=False)}


class Viewer:
    def __init__(self, writer, config, width, height, eye=(-1.5, 3, 3), center=(0, 1, 0), up=(0, 1, 0)):
        self.w = writer
        self.cfg = config
        self.width, self.height = width, height
        self.cam_main = Camera(CAMERA_FOV, float(width) / float(height), CAMERA_NEAR)
        self.cam_main.look_at(eye, center, up)             # OrbitController.cpp:14-16
        self.cam_depth = Camera(CAMERA_FOV, 1.0, CAMERA_NEAR)
        self.cam = self.cam_main
        self.fbo_main = self.tex_color_main = self.tex_depth_main = None
        self.main_ms = None
        self.fbo_shadow = self.tex_depth_shadow = None
        self.fxaa = None
        self.tex_color_fxaa = None
        self.program_cache = {}
        self.pipeline_cache = {}
        self.ibl_files = None        # optional dict(cube=path, irradiance=path, prefilter=path) => TEX_LOAD_RAW
        self.ibl_store = None        # optional dict(...) => TEX_STORE_RAW after generation
        # Viewer::create (Viewer.cpp:21-58)
        w = self.w
        self.block_scene = w.create_block("UniformsScene", 64)
        self.block_model = w.create_block("UniformsModel", 256)
        self.block_material = w.create_block("UniformsMaterial", 48)
        self.shadow_placeholder = self._texture2d_default(1, 1, T.TextureFormat_FLOAT32, T.TextureUsage_Sampler)
        self.ibl_placeholder = self._texture_cube_default(1, 1, T.TextureUsage_Sampler)

    # --- helpers (Viewer.cpp:848-892) -------------------------------------
    def _texture_cube_default(self, w, h, usage, mipmaps=False):
        t = self.w.create_texture(w, h, T.TextureType_CUBE, T.TextureFormat_RGBA8, usage, mipmaps, False)
        self.w.tex_set_sampler(t, T.Filter_LINEAR_MIPMAP_LINEAR if mipmaps else T.Filter_LINEAR, T.Filter_LINEAR)
        self.w.tex_init(t)
        return (t, T.TextureType_CUBE, T.TextureFormat_RGBA8)

    def _texture2d_default(self, w, h, fmt, usage, mipmaps=False):
        t = self.w.create_texture(w, h, T.TextureType_2D, fmt, usage, mipmaps, False)
        self.w.tex_set_sampler(t, T.Filter_LINEAR_MIPMAP_LINEAR if mipmaps else T.Filter_LINEAR, T.Filter_LINEAR)
        self.w.tex_init(t)
        return (t, T.TextureType_2D, fmt)

    # --- frame ----------------------------------------------------------------
    def draw_frame(self, scene):
        self.scene = scene
        self.cam_main.reverse_z = self.cfg.reverse_z          # ViewerSoftware::configRenderer
        self.cam_depth.reverse_z = self.cfg.reverse_z
        self._setup_main_buffers()
        self._setup_shadow_buffers()
        self._init_skybox_ibl()
        self._setup_scene()
        self._draw_shadow_map()
        self._fxaa_setup()
        self.w.begin_pass(self.fbo_main, True, self.cfg.depth_test, self.cfg.clear_color,
                          0.0 if self.cfg.reverse_z else 1.0)
        self.w.viewport(0, 0, self.width, self.height)
        self._draw_scene(False)
        self.w.end_pass()
        if self.cfg.aa_type == AA_FXAA:
            self._fxaa_draw()

    def _setup_main_buffers(self):
        ms = self.cfg.aa_type == AA_MSAA
        w = self.w
        if self.tex_color_main is None or self.main_ms != ms:
            t = w.create_texture(self.width, self.height, T.TextureType_2D, T.TextureFormat_RGBA8,
                                 T.TextureUsage_AttachmentColor | T.TextureUsage_RendererOutput, False, ms)
            w.tex_set_sampler(t, T.Filter_LINEAR, T.Filter_LINEAR)
            w.tex_init(t)
            self.tex_color_main = t
            d = w.create_texture(self.width, self.height, T.TextureType_2D, T.TextureFormat_FLOAT32,
                                 T.TextureUsage_AttachmentDepth, False, ms)
            w.tex_set_sampler(d, T.Filter_NEAREST, T.Filter_NEAREST)
            w.tex_init(d)
            self.tex_depth_main = d
            self.main_ms = ms
        if self.fbo_main is None:
            self.fbo_main = w.create_fbo(False)
        w.fbo_color(self.fbo_main, self.tex_color_main, 0)
        w.fbo_depth(self.fbo_main, self.tex_depth_main)
        w.fbo_offscreen(self.fbo_main, False)

    def _setup_shadow_buffers(self):
        if not self.cfg.shadow_map:
            return
        w = self.w
        if self.fbo_shadow is None:
            self.fbo_shadow = w.create_fbo(True)
        if self.tex_depth_shadow is None:
            t = w.create_texture(SHADOW_MAP_SIZE, SHADOW_MAP_SIZE, T.TextureType_2D, T.TextureFormat_FLOAT32,
                                 T.TextureUsage_Sampler | T.TextureUsage_AttachmentDepth, False, False)
            w.tex_set_sampler(t, T.Filter_NEAREST, T.Filter_NEAREST, T.Wrap_CLAMP_TO_BORDER,
                              T.Wrap_CLAMP_TO_BORDER, T.Wrap_CLAMP_TO_EDGE,
                              T.Border_BLACK if self.cfg.reverse_z else T.Border_WHITE)
            w.tex_init(t)
            w.fbo_depth(self.fbo_shadow, t)
            self.tex_depth_shadow = (t, T.TextureType_2D, T.TextureFormat_FLOAT32)

    # --- IBL (Viewer.cpp:718-809, Environment.cpp) -------------------------------
    def _ibl_enabled(self):
        return self.cfg.show_skybox and self.cfg.pbr_ibl and self.scene.skybox.ibl_ready

    def _init_skybox_ibl(self):
        if not (self.cfg.show_skybox and self.cfg.pbr_ibl):
            return
        sky = self.scene.skybox
        if sky.ibl_ready:
            return
        mat = sky.material
        if not mat.textures:
            self._setup_textures(mat)
        w = self.w
        if T.MaterialTexType_CUBE not in mat.textures:
            eq = mat.textures[T.MaterialTexType_EQUIRECTANGULAR]
            d = w.tex_desc[eq[0]]
            size = min(d["width"], d["height"])
            cvt = self._texture_cube_default(size, size, T.TextureUsage_AttachmentColor | T.TextureUsage_Sampler)
            if self.ibl_files:
                w.tex_load_raw(cvt[0], self.ibl_files["cube"])
            else:
                ctx = self._cube_ctx(T.Shading_Skybox, eq, T.MaterialTexType_EQUIRECTANGULAR)
                self._draw_cube_faces(ctx, size, size, cvt[0], 0)
            if self.ibl_store:
                w.tex_store_raw(cvt[0], self.ibl_store["cube"])
            mat.textures[T.MaterialTexType_CUBE] = cvt
            w.wait_idle()
            del mat.textures[T.MaterialTexType_EQUIRECTANGULAR]
            mat.defines = self._defines(mat)
            mat.obj = None
        cube = mat.textures[T.MaterialTexType_CUBE]
        cube_w = w.tex_desc[cube[0]]["width"]
        irr = self._texture_cube_default(IRRADIANCE_SIZE, IRRADIANCE_SIZE,
                                         T.TextureUsage_AttachmentColor | T.TextureUsage_Sampler)
        if self.ibl_files:
            w.tex_load_raw(irr[0], self.ibl_files["irradiance"])
        else:
            ctx = self._cube_ctx(T.Shading_IBL_Irradiance, cube, T.MaterialTexType_CUBE)
            self._draw_cube_faces(ctx, IRRADIANCE_SIZE, IRRADIANCE_SIZE, irr[0], 0)
        if self.ibl_store:
            w.tex_store_raw(irr[0], self.ibl_store["irradiance"])
        mat.textures[T.MaterialTexType_IBL_IRRADIANCE] = irr
        pre = self._texture_cube_default(PREFILTER_SIZE, PREFILTER_SIZE,
                                         T.TextureUsage_AttachmentColor | T.TextureUsage_Sampler, True)
        if self.ibl_files:
            w.tex_load_raw(pre[0], self.ibl_files["prefilter"])
        else:
            ctx = self._cube_ctx(T.Shading_IBL_Prefilter, cube, T.MaterialTexType_CUBE)
            blk = w.create_block("UniformsPrefilter", 8)
            ctx["blocks"][T.UniformBlock_IBLPrefilter] = blk
            for level in range(PREFILTER_LEVELS):
                lw = max(1, PREFILTER_SIZE >> level)
                data = struct.pack("<2f", float(cube_w), float(level) / float(PREFILTER_LEVELS - 1))
                self._draw_cube_faces(ctx, lw, lw, pre[0], level, lambda: w.block_data(blk, data))
        if self.ibl_store:
            w.tex_store_raw(pre[0], self.ibl_store["prefilter"])
        mat.textures[T.MaterialTexType_IBL_PREFILTER] = pre
        w.wait_idle()
        sky.ibl_ready = True

    def _cube_ctx(self, shading, tex_in, tex_type):
        w = self.w
        cv, ci = cube_mesh()
        ctx = dict(fbo=w.create_fbo(True), vao=w.create_vao(cv, ci))
        ctx["program"] = w.create_program(shading, [_SAMPLER_DEFINE[tex_type]])
        s = w.create_sampler(_SAMPLER_NAME[tex_type], tex_in[1], tex_in[2])
        w.sampler_tex(s, tex_in[0])
        ctx["samplers"] = {tex_type: s}
        ctx["block_model"] = w.create_block("UniformsModel", 256)
        ctx["blocks"] = {T.UniformBlock_Model: ctx["block_model"]}
        ctx["pipeline"] = w.create_pipeline(T.RenderStates())
        return ctx

    def _draw_cube_faces(self, ctx, width, height, tex_out, level, before_draw=None):
        views = [((1, 0, 0), (0, -1, 0)), ((-1, 0, 0), (0, -1, 0)), ((0, 1, 0), (0, 0, 1)),
                 ((0, -1, 0), (0, 0, -1)), ((0, 0, 1), (0, -1, 0)), ((0, 0, -1), (0, -1, 0))]
        cam = Camera(90.0, 1.0, 0.1)
        w = self.w
        zero4 = np.zeros((4, 4), np.float32)
        for i, (center, up) in enumerate(views):
            cam.look_at((0, 0, 0), center, up)
            v = np.eye(4, dtype=np.float32)
            v[:3, :3] = cam.view()[:3, :3]
            mvp = cam.projection() @ v
            w.block_data(ctx["block_model"], pack_uniforms_model(False, zero4, mvp, np.zeros((3, 3)), zero4))
            if before_draw:
                before_draw()
            w.fbo_color(ctx["fbo"], tex_out, level, i)
            w.begin_pass(ctx["fbo"], True, False, (0, 0, 0, 0), 1.0)
            w.viewport(0, 0, width, height)
            w.draw(ctx["vao"], ctx["program"], ctx["pipeline"], ctx["blocks"], ctx["samplers"])
            w.end_pass()

    # --- scene setup (Viewer.cpp:215-291, 374-386, 498-671) ------------------------
    def _setup_scene(self):
        cfg, sc = self.cfg, self.scene
        MB, SB, TB = T.UniformBlock_Model, T.UniformBlock_Scene, T.UniformBlock_Material
        if cfg.show_light:
            self._pipeline_setup(sc.point_light, sc.point_light.material.shading, (MB, TB))
        if cfg.world_axis:
            self._pipeline_setup(sc.world_axis, sc.world_axis.material.shading, (MB, TB))
        if cfg.show_floor:
            self._setup_mesh(sc.floor)
        if cfg.show_skybox:
            def sky_states(rs):
                rs.depthFunc = T.DepthFunc_GEQUAL if cfg.reverse_z else T.DepthFunc_LEQUAL
                rs.depthMask = False
            self._pipeline_setup(sc.skybox, sc.skybox.material.shading, (MB,), sky_states)
        self._walk(sc.model.root, lambda mesh: self._setup_mesh(sc.drawable_for(mesh)))

    def _walk(self, node, fn):
        for m in node.meshes:
            fn(m)
        for c in node.children:
            self._walk(c, fn)

    def _setup_mesh(self, d):
        MB, SB, TB = T.UniformBlock_Model, T.UniformBlock_Scene, T.UniformBlock_Material
        if self.cfg.wireframe:
            def wf(rs):
                rs.polygonMode = T.PolygonMode_LINE
            self._pipeline_setup(d, T.Shading_BaseColor, (MB, SB, TB), wf)
        else:
            self._pipeline_setup(d, d.material.shading, (MB, SB, TB))

    def _pipeline_setup(self, d, shading, blocks, extra=None):
        w = self.w
        if d.vao is None:
            d.vao = w.create_vao(d.vertices, d.indices)
        mat = d.material
        if mat.obj is not None and mat.obj["shading"] != shading:
            mat.obj = None
        if not mat.textures:
            self._setup_textures(mat)
            mat.defines = self._defines(mat)
        if mat.obj is None:
            key = (shading, tuple(sorted(mat.defines)))
            if key not in self.program_cache:
                self.program_cache[key] = w.create_program(shading, sorted(mat.defines))
            obj = dict(shading=shading, program=self.program_cache[key], samplers={}, blocks={}, sampler_tex={})
            for k, tex in mat.textures.items():
                if k in _SAMPLER_NAME:
                    s = w.create_sampler(_SAMPLER_NAME[k], tex[1], tex[2])
                    w.sampler_tex(s, tex[0])
                    obj["samplers"][k] = s
                    obj["sampler_tex"][k] = tex[0]
            for b in blocks:
                obj["blocks"][b] = {T.UniformBlock_Scene: self.block_scene, T.UniformBlock_Model: self.block_model,
                                    T.UniformBlock_Material: self.block_material}[b]
            mat.obj = obj
        rs = T.RenderStates()
        rs.blend = mat.alpha_blend
        rs.set_blend_factor(T.BlendFactor_SRC_ALPHA, T.BlendFactor_ONE_MINUS_SRC_ALPHA)
        rs.depthTest = self.cfg.depth_test
        rs.depthMask = not rs.blend
        rs.depthFunc = T.DepthFunc_GREATER if self.cfg.reverse_z else T.DepthFunc_LESS
        rs.cullFace = self.cfg.cull_face and not mat.double_sided
        rs.primitiveType = d.primitive
        rs.polygonMode = T.PolygonMode_FILL
        rs.lineWidth = mat.line_width
        if extra:
            extra(rs)
        pk = (shading,) + rs.key()
        if pk not in self.pipeline_cache:
            self.pipeline_cache[pk] = w.create_pipeline(rs)
        mat.obj["pipeline"] = self.pipeline_cache[pk]

    def _setup_textures(self, mat):
        w = self.w
        for k, td in mat.texture_data.items():
            if k in (T.MaterialTexType_IBL_IRRADIANCE, T.MaterialTexType_IBL_PREFILTER):
                continue
            img0 = td["images"][0]
            h, wd = img0.shape[0], img0.shape[1]
            cube = k == T.MaterialTexType_CUBE
            mip = (not cube) and self.cfg.mipmaps
            t = w.create_texture(wd, h, T.TextureType_CUBE if cube else T.TextureType_2D, T.TextureFormat_RGBA8,
                                 T.TextureUsage_Sampler | T.TextureUsage_UploadData, mip, False)
            w.tex_set_sampler(t, T.Filter_LINEAR_MIPMAP_LINEAR if mip else T.Filter_LINEAR, T.Filter_LINEAR,
                              td["wrap_u"], td["wrap_v"], td["wrap_w"] if cube else T.Wrap_CLAMP_TO_EDGE)
            w.tex_set_data(t, td["images"])
            mat.textures[k] = (t, T.TextureType_CUBE if cube else T.TextureType_2D, T.TextureFormat_RGBA8)
        if mat.shading != T.Shading_Skybox:
            mat.textures[T.MaterialTexType_SHADOWMAP] = self.shadow_placeholder
        if mat.shading == T.Shading_PBR:
            mat.textures[T.MaterialTexType_IBL_IRRADIANCE] = self.ibl_placeholder
            mat.textures[T.MaterialTexType_IBL_PREFILTER] = self.ibl_placeholder

    @staticmethod
    def _defines(mat):
        return {_SAMPLER_DEFINE[k] for k in mat.textures if k in _SAMPLER_DEFINE}

    # --- draw (Viewer.cpp:143-168, 293-396, 673-716) -------------------------------------
    def _draw_shadow_map(self):
        if not self.cfg.shadow_map:
            return
        w = self.w
        w.begin_pass(self.fbo_shadow, False, True, (0, 0, 0, 0), 0.0 if self.cfg.reverse_z else 1.0)
        w.viewport(0, 0, SHADOW_MAP_SIZE, SHADOW_MAP_SIZE)
        self.cam_depth.look_at(self.cfg.point_light_position, (0, 0, 0), (0, 1, 0))
        self.cam = self.cam_depth
        self._draw_scene(True)
        w.end_pass()
        self.cam = self.cam_main

    def _update_uniform_scene(self):
        c = self.cfg
        self.w.block_data(self.block_scene, pack_uniforms_scene(c.ambient_color, self.cam.eye,
                                                                c.point_light_position, c.point_light_color))

    def _update_uniform_model(self, model, view):
        mvp = self.cam.projection() @ view @ model
        it = np.linalg.inv(model.astype(np.float64)).T[:3, :3].astype(np.float32)
        shadow = np.zeros((4, 4), np.float32)
        if self.cfg.shadow_map:
            bias = np.array([[0.5, 0, 0, 0.5], [0, 0.5, 0, 0.5], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
            shadow = bias @ self.cam_depth.projection() @ self.cam_depth.view() @ model
        self.w.block_data(self.block_model, pack_uniforms_model(self.cfg.reverse_z, model, mvp, it, shadow))

    def _update_uniform_material(self, mat, specular=1.0):
        c = self.cfg
        self.w.block_data(self.block_material, pack_uniforms_material(
            c.show_light, self._ibl_enabled(), c.shadow_map, mat.point_size, specular, mat.base_color))

    def _pipeline_draw(self, d):
        o = d.material.obj
        self.w.draw(d.vao, o["program"], o["pipeline"], o["blocks"], o["samplers"])

    def _set_sampler(self, obj, key, tex):
        if key in obj["samplers"] and obj["sampler_tex"].get(key) != tex[0]:
            self.w.sampler_tex(obj["samplers"][key], tex[0])
            obj["sampler_tex"][key] = tex[0]

    def _draw_mesh(self, d, shadow_pass, specular):
        mat = d.material
        self._update_uniform_material(mat, specular)
        if mat.shading == T.Shading_PBR:
            if self._ibl_enabled():
                st = self.scene.skybox.material.textures
                self._set_sampler(mat.obj, T.MaterialTexType_IBL_IRRADIANCE, st[T.MaterialTexType_IBL_IRRADIANCE])
                self._set_sampler(mat.obj, T.MaterialTexType_IBL_PREFILTER, st[T.MaterialTexType_IBL_PREFILTER])
            else:
                self._set_sampler(mat.obj, T.MaterialTexType_IBL_IRRADIANCE, self.ibl_placeholder)
                self._set_sampler(mat.obj, T.MaterialTexType_IBL_PREFILTER, self.ibl_placeholder)
        if self.cfg.shadow_map:
            self._set_sampler(mat.obj, T.MaterialTexType_SHADOWMAP,
                              self.shadow_placeholder if shadow_pass else self.tex_depth_shadow)
        self._pipeline_draw(d)

    def _draw_nodes(self, node, shadow_pass, transform, blend, specular=1.0):
        model = (transform @ node.transform).astype(np.float32)
        self._update_uniform_model(model, self.cam.view())
        for mesh in node.meshes:
            d = self.scene.drawable_for(mesh)
            if d.material.alpha_blend != blend:
                continue
            self._draw_mesh(d, shadow_pass, specular)
        for c in node.children:
            self._draw_nodes(c, shadow_pass, model, blend, specular)

    def _draw_scene(self, shadow_pass):
        cfg, sc = self.cfg, self.scene
        self._update_uniform_scene()
        self._update_uniform_model(np.eye(4, dtype=np.float32), self.cam.view())
        if not shadow_pass and cfg.show_light:
            self._update_uniform_material(sc.point_light.material)
            self._pipeline_draw(sc.point_light)
        if not shadow_pass and cfg.world_axis:
            self._update_uniform_material(sc.world_axis.material)
            self._pipeline_draw(sc.world_axis)
        if not shadow_pass and cfg.show_floor:
            self._draw_mesh(sc.floor, shadow_pass, 0.0)
        self._draw_nodes(sc.model.root, shadow_pass, sc.model.centered, False)
        if not shadow_pass and cfg.show_skybox:
            v = np.eye(4, dtype=np.float32)
            v[:3, :3] = self.cam.view()[:3, :3]
            self._update_uniform_model(np.eye(4, dtype=np.float32), v)
            self._pipeline_draw(sc.skybox)
        self._draw_nodes(sc.model.root, shadow_pass, sc.model.centered, True)

    # --- FXAA (Viewer.cpp:170-213, QuadFilter.cpp) --------------------------------------
    def _fxaa_setup(self):
        if self.cfg.aa_type != AA_FXAA:
            return
        w = self.w
        if self.tex_color_fxaa is None:
            t = w.create_texture(self.width, self.height, T.TextureType_2D, T.TextureFormat_RGBA8,
                                 T.TextureUsage_Sampler | T.TextureUsage_AttachmentColor, False, False)
            w.tex_set_sampler(t, T.Filter_LINEAR, T.Filter_LINEAR)
            w.tex_init(t)
            self.tex_color_fxaa = t
        if self.fxaa is None:
            pos = np.array([(1, -1, 0), (-1, -1, 0), (1, 1, 0), (-1, 1, 0)], np.float32)
            uv = np.array([(1, 0), (0, 0), (1, 1), (0, 1)], np.float32)
            f = dict(fbo=w.create_fbo(False),
                     vao=w.create_vao(A.make_vertices(pos, uv), np.array([0, 1, 2, 1, 2, 3], np.int32)),
                     program=w.create_program(T.Shading_FXAA, []))
            f["sampler"] = w.create_sampler("u_screenTexture", T.TextureType_2D, T.TextureFormat_RGBA8)
            f["block"] = w.create_block("UniformsQuadFilter", 8)
            w.block_data(f["block"], struct.pack("<2f", 0.0, 0.0))
            f["pipeline"] = w.create_pipeline(T.RenderStates())
            self.fxaa = f
        w.fbo_color(self.fbo_main, self.tex_color_fxaa, 0)
        w.fbo_offscreen(self.fbo_main, True)
        f = self.fxaa
        w.block_data(f["block"], struct.pack("<2f", float(self.width), float(self.height)))
        w.sampler_tex(f["sampler"], self.tex_color_fxaa)
        w.fbo_color(f["fbo"], self.tex_color_main, 0)

    def _fxaa_draw(self):
        w, f = self.w, self.fxaa
        w.begin_pass(f["fbo"], True, False, (0, 0, 0, 0), 1.0)
        w.viewport(0, 0, self.width, self.height)
        w.draw(f["vao"], f["program"], f["pipeline"], {T.UniformBlock_QuadFilter: f["block"]},
               {T.MaterialTexType_QUAD_FILTER: f["sampler"]})
        w.end_pass()

    # --- outputs -------------------------------------------------------------------
    def readback_all(self):
        w = self.w
        w.wait_idle()
        w.readback(self.tex_color_main, "color")
        w.readback(self.tex_depth_main, "depth")
        if self.cfg.shadow_map and self.tex_depth_shadow is not None:
            w.readback(self.tex_depth_shadow[0], "shadow")
        if self.tex_color_fxaa is not None:
            w.readback(self.tex_color_fxaa, "color_prefxaa")
