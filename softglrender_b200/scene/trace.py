"""Renderer-API command traces ("scene packs").

A trace is a little-endian stream of the calls a Viewer makes on the abstract
``Renderer`` interface (reference: src/Render/Renderer.h:24-59 and the companion
objects in src/Render/{Texture,Framebuffer,Uniform,ShaderProgram,Vertex}.h).  The
same trace is replayed by three players -- the compiled reference RendererSoft
(oracle/_ref), the scalar CPU restatement (oracle/) and RendererCUDA -- so that all
of them see byte-identical inputs.

Layout: magic "SGLT", u32 version, then commands ``u32 opcode, u32 nbytes, payload``.
"""
import struct
import numpy as np

MAGIC = b"SGLT"
VERSION = 1

# opcodes (mirrored in softglrender_b200/harness/trace_format.h)
OP_CREATE_TEXTURE = 1
OP_TEX_SET_SAMPLER = 2
OP_TEX_INIT = 3
OP_TEX_SET_DATA = 4
OP_TEX_LOAD_RAW = 6
OP_TEX_STORE_RAW = 7
OP_CREATE_VAO = 10
OP_VAO_UPDATE = 11
OP_CREATE_PROGRAM = 12
OP_CREATE_BLOCK = 13
OP_CREATE_SAMPLER = 14
OP_CREATE_PIPELINE = 15
OP_CREATE_FBO = 16
OP_FBO_COLOR = 20
OP_FBO_DEPTH = 21
OP_FBO_OFFSCREEN = 22
OP_BEGIN_PASS = 30
OP_VIEWPORT = 31
OP_BLOCK_DATA = 32
OP_SAMPLER_TEX = 33
OP_DRAW = 34
OP_END_PASS = 35
OP_WAIT_IDLE = 36
OP_READBACK = 40
OP_FRAME_BEGIN = 50
OP_FRAME_END = 51

# enums of the reference API (src/Render/Texture.h:17-77, RenderStates.h:13-99)
Wrap_REPEAT, Wrap_MIRRORED_REPEAT, Wrap_CLAMP_TO_EDGE, Wrap_CLAMP_TO_BORDER = range(4)
(Filter_NEAREST, Filter_LINEAR, Filter_NEAREST_MIPMAP_NEAREST, Filter_LINEAR_MIPMAP_NEAREST,
 Filter_NEAREST_MIPMAP_LINEAR, Filter_LINEAR_MIPMAP_LINEAR) = range(6)
Border_BLACK, Border_WHITE = 0, 1
TextureType_2D, TextureType_CUBE = 0, 1
TextureFormat_RGBA8, TextureFormat_FLOAT32 = 0, 1
TextureUsage_Sampler = 1
TextureUsage_UploadData = 2
TextureUsage_AttachmentColor = 4
TextureUsage_AttachmentDepth = 8
TextureUsage_RendererOutput = 16
(DepthFunc_NEVER, DepthFunc_LESS, DepthFunc_EQUAL, DepthFunc_LEQUAL, DepthFunc_GREATER,
 DepthFunc_NOTEQUAL, DepthFunc_GEQUAL, DepthFunc_ALWAYS) = range(8)
(BlendFactor_ZERO, BlendFactor_ONE, BlendFactor_SRC_COLOR, BlendFactor_SRC_ALPHA,
 BlendFactor_DST_COLOR, BlendFactor_DST_ALPHA, BlendFactor_ONE_MINUS_SRC_COLOR,
 BlendFactor_ONE_MINUS_SRC_ALPHA, BlendFactor_ONE_MINUS_DST_COLOR,
 BlendFactor_ONE_MINUS_DST_ALPHA) = range(10)
BlendFunc_ADD, BlendFunc_SUBTRACT, BlendFunc_REVERSE_SUBTRACT, BlendFunc_MIN, BlendFunc_MAX = range(5)
PolygonMode_POINT, PolygonMode_LINE, PolygonMode_FILL = range(3)
Primitive_POINT, Primitive_LINE, Primitive_TRIANGLE = range(3)

# Viewer-level enums (src/Viewer/Material.h:20-62) -- only used as map keys / program ids
(Shading_Unknown, Shading_BaseColor, Shading_BlinnPhong, Shading_PBR, Shading_Skybox,
 Shading_IBL_Irradiance, Shading_IBL_Prefilter, Shading_FXAA) = range(8)
(MaterialTexType_NONE, MaterialTexType_ALBEDO, MaterialTexType_NORMAL, MaterialTexType_EMISSIVE,
 MaterialTexType_AMBIENT_OCCLUSION, MaterialTexType_METAL_ROUGHNESS, MaterialTexType_CUBE,
 MaterialTexType_EQUIRECTANGULAR, MaterialTexType_IBL_IRRADIANCE, MaterialTexType_IBL_PREFILTER,
 MaterialTexType_QUAD_FILTER, MaterialTexType_SHADOWMAP) = range(12)
(UniformBlock_Scene, UniformBlock_Model, UniformBlock_Material, UniformBlock_QuadFilter,
 UniformBlock_IBLPrefilter) = range(5)


class RenderStates:
    """POD twin of SoftGL::RenderStates (src/Render/RenderStates.h:80-92)."""

    def __init__(self):
        self.blend = False
        self.blendFuncRgb = BlendFunc_ADD
        self.blendSrcRgb = BlendFactor_ONE
        self.blendDstRgb = BlendFactor_ZERO
        self.blendFuncAlpha = BlendFunc_ADD
        self.blendSrcAlpha = BlendFactor_ONE
        self.blendDstAlpha = BlendFactor_ZERO
        self.depthTest = False
        self.depthMask = True
        self.depthFunc = DepthFunc_LESS
        self.cullFace = False
        self.primitiveType = Primitive_TRIANGLE
        self.polygonMode = PolygonMode_FILL
        self.lineWidth = 1.0

    def set_blend_factor(self, src, dst):
        self.blendSrcRgb = self.blendSrcAlpha = src
        self.blendDstRgb = self.blendDstAlpha = dst

    def key(self):
        return (self.blend, self.blendFuncRgb, self.blendSrcRgb, self.blendDstRgb, self.blendFuncAlpha,
                self.blendSrcAlpha, self.blendDstAlpha, self.depthTest, self.depthMask, self.depthFunc,
                self.cullFace, self.primitiveType, self.polygonMode, float(self.lineWidth))

    def pack(self):
        return struct.pack("<13if", int(self.blend), self.blendFuncRgb, self.blendSrcRgb, self.blendDstRgb,
                           self.blendFuncAlpha, self.blendSrcAlpha, self.blendDstAlpha, int(self.depthTest),
                           int(self.depthMask), self.depthFunc, int(self.cullFace), self.primitiveType,
                           self.polygonMode, float(self.lineWidth))


def _s(text):
    b = text.encode("utf-8")
    return struct.pack("<I", len(b)) + b


class TraceWriter:
    """Records Renderer-API calls; object ids are per-type creation indices."""

    def __init__(self):
        self.chunks = [MAGIC, struct.pack("<I", VERSION)]
        self.n = {"tex": 0, "vao": 0, "prog": 0, "block": 0, "sampler": 0, "pipe": 0, "fbo": 0}
        self.tex_desc = {}

    def _cmd(self, op, payload=b""):
        self.chunks.append(struct.pack("<II", op, len(payload)))
        if payload:
            self.chunks.append(payload)

    def _new(self, kind):
        i = self.n[kind]
        self.n[kind] += 1
        return i

    # --- resources -------------------------------------------------------
    def create_texture(self, width, height, type=TextureType_2D, format=TextureFormat_RGBA8,
                       usage=TextureUsage_Sampler, use_mipmaps=False, multi_sample=False):
        self._cmd(OP_CREATE_TEXTURE, struct.pack("<7i", width, height, type, format, usage,
                                                 int(use_mipmaps), int(multi_sample)))
        t = self._new("tex")
        self.tex_desc[t] = dict(width=width, height=height, type=type, format=format, usage=usage,
                                use_mipmaps=use_mipmaps, multi_sample=multi_sample)
        return t

    def tex_set_sampler(self, tex, filter_min=Filter_NEAREST, filter_mag=Filter_NEAREST,
                        wrap_s=Wrap_CLAMP_TO_EDGE, wrap_t=Wrap_CLAMP_TO_EDGE, wrap_r=Wrap_CLAMP_TO_EDGE,
                        border=Border_BLACK):
        self._cmd(OP_TEX_SET_SAMPLER, struct.pack("<7i", tex, filter_min, filter_mag, wrap_s, wrap_t,
                                                  wrap_r, border))

    def tex_init(self, tex):
        self._cmd(OP_TEX_INIT, struct.pack("<i", tex))

    def tex_set_data(self, tex, buffers):
        """buffers: list of (H,W,4) uint8 arrays (RGBA8) or (H,W) float32 arrays, one per layer."""
        parts = [struct.pack("<ii", tex, len(buffers))]
        for b in buffers:
            b = np.ascontiguousarray(b)
            h, w = b.shape[0], b.shape[1]
            raw = b.tobytes()
            assert len(raw) == w * h * 4
            parts.append(struct.pack("<II", w, h))
            parts.append(raw)
        self._cmd(OP_TEX_SET_DATA, b"".join(parts))

    def tex_load_raw(self, tex, path):
        self._cmd(OP_TEX_LOAD_RAW, struct.pack("<i", tex) + _s(path))

    def tex_store_raw(self, tex, path):
        self._cmd(OP_TEX_STORE_RAW, struct.pack("<i", tex) + _s(path))

    def create_vao(self, vertices, indices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).tobytes()
        i = np.ascontiguousarray(indices, dtype=np.int32).tobytes()
        assert len(v) % 64 == 0
        self._cmd(OP_CREATE_VAO, struct.pack("<I", len(v)) + v + struct.pack("<I", len(i)) + i)
        return self._new("vao")

    def vao_update(self, vao, vertices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).tobytes()
        self._cmd(OP_VAO_UPDATE, struct.pack("<iI", vao, len(v)) + v)

    def create_program(self, shading_model, defines=()):
        p = struct.pack("<iI", shading_model, len(defines)) + b"".join(_s(d) for d in sorted(defines))
        self._cmd(OP_CREATE_PROGRAM, p)
        return self._new("prog")

    def create_block(self, name, size):
        self._cmd(OP_CREATE_BLOCK, _s(name) + struct.pack("<i", size))
        return self._new("block")

    def create_sampler(self, name, tex_type, tex_format):
        self._cmd(OP_CREATE_SAMPLER, _s(name) + struct.pack("<ii", tex_type, tex_format))
        return self._new("sampler")

    def create_pipeline(self, rs):
        self._cmd(OP_CREATE_PIPELINE, rs.pack())
        return self._new("pipe")

    def create_fbo(self, offscreen):
        self._cmd(OP_CREATE_FBO, struct.pack("<i", int(offscreen)))
        return self._new("fbo")

    # --- state / pass ----------------------------------------------------
    def fbo_color(self, fbo, tex, level=0, face=-1):
        self._cmd(OP_FBO_COLOR, struct.pack("<4i", fbo, tex, face, level))

    def fbo_depth(self, fbo, tex):
        self._cmd(OP_FBO_DEPTH, struct.pack("<2i", fbo, tex))

    def fbo_offscreen(self, fbo, flag):
        self._cmd(OP_FBO_OFFSCREEN, struct.pack("<2i", fbo, int(flag)))

    def begin_pass(self, fbo, color_flag=False, depth_flag=False, clear_color=(0, 0, 0, 0), clear_depth=1.0):
        self._cmd(OP_BEGIN_PASS, struct.pack("<3i5f", fbo, int(color_flag), int(depth_flag),
                                             *[float(c) for c in clear_color], float(clear_depth)))

    def viewport(self, x, y, w, h):
        self._cmd(OP_VIEWPORT, struct.pack("<4i", x, y, w, h))

    def block_data(self, block, data, offset=0):
        data = bytes(data)
        self._cmd(OP_BLOCK_DATA, struct.pack("<3i", block, offset, len(data)) + data)

    def sampler_tex(self, sampler, tex):
        self._cmd(OP_SAMPLER_TEX, struct.pack("<2i", sampler, tex))

    def draw(self, vao, program, pipeline, blocks, samplers):
        """blocks/samplers: dict key -> id (the ShaderResources maps, src/Render/Uniform.h:60-64)."""
        p = [struct.pack("<4i", vao, program, pipeline, len(blocks))]
        for k in sorted(blocks):
            p.append(struct.pack("<2i", k, blocks[k]))
        p.append(struct.pack("<i", len(samplers)))
        for k in sorted(samplers):
            p.append(struct.pack("<2i", k, samplers[k]))
        self._cmd(OP_DRAW, b"".join(p))

    def end_pass(self):
        self._cmd(OP_END_PASS)

    def wait_idle(self):
        self._cmd(OP_WAIT_IDLE)

    def readback(self, tex, tag, layer=0, level=0):
        self._cmd(OP_READBACK, struct.pack("<3i", tex, layer, level) + _s(tag))

    def frame_begin(self):
        self._cmd(OP_FRAME_BEGIN)

    def frame_end(self):
        self._cmd(OP_FRAME_END)

    def tobytes(self):
        return b"".join(self.chunks)

    def save(self, path):
        with open(path, "wb") as f:
            for c in self.chunks:
                f.write(c)


def read_outputs(path):
    """Parse a player's output file: {tag: ndarray}.  Records: str tag, i32 w,h,format,samples, u32 nbytes, data."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        (n,) = struct.unpack_from("<I", data, off)
        off += 4
        tag = data[off:off + n].decode()
        off += n
        w, h, fmt, samples, nbytes = struct.unpack_from("<4iI", data, off)
        off += 20
        raw = data[off:off + nbytes]
        off += nbytes
        if fmt == TextureFormat_RGBA8:
            a = np.frombuffer(raw, dtype=np.uint8)
            a = a.reshape(h, w, samples, 4) if samples > 1 else a.reshape(h, w, 4)
        else:
            a = np.frombuffer(raw, dtype=np.float32)
            a = a.reshape(h, w, samples) if samples > 1 else a.reshape(h, w)
        out[tag] = a
    return out
