"""Host-side asset loading for the headless harness (glTF 2.0, Wavefront OBJ, images).

The reference does this with assimp + stb (src/Viewer/ModelLoader.cpp:185-505); that code
is host-side input preparation and out of scope for the accelerated path, so this is an
independent loader that produces the same *shape* of data the Renderer API consumes:
64-byte vertices {vec3 pos@0, vec2 uv@16, vec3 normal@32, vec3 tangent@48}
(src/Viewer/Model.h:20-25), int32 indices and decoded RGBA8 images.  Every renderer
under comparison is fed the arrays produced here, so parity does not depend on it.
"""
import json
import os
import struct
import numpy as np

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def load_image_rgba(path):
    """Decode to (H,W,4) uint8, channel expansion as ImageUtils::readImageRGBA (src/Base/ImageUtils.cpp:19-70)."""
    from PIL import Image
    img = Image.open(path)
    if img.mode in ("L", "1", "I", "I;16"):
        img = img.convert("L")
        a = np.asarray(img, dtype=np.uint8)
        out = np.empty(a.shape + (4,), np.uint8)
        out[..., 0] = out[..., 1] = out[..., 2] = a
        out[..., 3] = 255
        return out
    if img.mode == "LA":
        a = np.asarray(img, dtype=np.uint8)
        out = np.empty(a.shape[:2] + (4,), np.uint8)
        out[..., 0] = out[..., 1] = out[..., 2] = a[..., 0]
        out[..., 3] = a[..., 1]
        return out
    return np.ascontiguousarray(np.asarray(img.convert("RGBA"), dtype=np.uint8))


def make_vertices(pos, uv=None, normal=None, tangent=None):
    n = len(pos)
    v = np.zeros((n, 16), np.float32)
    v[:, 0:3] = pos
    if uv is not None:
        v[:, 4:6] = uv
    if normal is not None:
        v[:, 8:11] = normal
    if tangent is not None:
        v[:, 12:15] = tangent
    return v


def calc_tangents(pos, uv, normal, indices):
    """Per-vertex tangents from UV gradients (role of aiProcess_CalcTangentSpace, ModelLoader.cpp:198-202)."""
    pos = np.asarray(pos, np.float64)
    uv = np.asarray(uv, np.float64)
    nrm = np.asarray(normal, np.float64)
    tri = np.asarray(indices, np.int64).reshape(-1, 3)
    p0, p1, p2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    t0, t1, t2 = uv[tri[:, 0]], uv[tri[:, 1]], uv[tri[:, 2]]
    v, w = p1 - p0, p2 - p0
    sx, sy = t1[:, 0] - t0[:, 0], t1[:, 1] - t0[:, 1]
    tx, ty = t2[:, 0] - t0[:, 0], t2[:, 1] - t0[:, 1]
    dirc = np.where(tx * sy - ty * sx < 0, -1.0, 1.0)
    deg = (sx * ty == sy * tx)
    sy = np.where(deg, 1.0, sy)
    ty = np.where(deg, 0.0, ty)
    tang = (w * sy[:, None] - v * ty[:, None]) * dirc[:, None]
    acc = np.zeros_like(pos)
    for k in range(3):
        np.add.at(acc, tri[:, k], tang)
    acc = acc - nrm * np.sum(acc * nrm, axis=1, keepdims=True)
    ln = np.linalg.norm(acc, axis=1, keepdims=True)
    fallback = np.cross(nrm, np.array([0.0, 1.0, 0.0]))
    fl = np.linalg.norm(fallback, axis=1, keepdims=True)
    fallback = np.where(fl > 1e-6, fallback / np.maximum(fl, 1e-30), np.array([1.0, 0.0, 0.0]))
    out = np.where(ln > 1e-12, acc / np.maximum(ln, 1e-30), fallback)
    return out.astype(np.float32)


class Mesh:
    def __init__(self):
        self.vertices = None       # (N,16) float32
        self.indices = None        # int32
        self.aabb = None           # (min3, max3)
        self.shading = "pbr"       # "pbr" | "blinnphong"
        self.alpha_blend = False
        self.double_sided = False
        self.base_color = (1.0, 1.0, 1.0, 1.0)
        self.textures = {}         # tex-type name -> dict(path, image, wrap_u, wrap_v)


class Node:
    def __init__(self):
        self.transform = np.eye(4, dtype=np.float32)
        self.meshes = []
        self.children = []


class Model:
    def __init__(self):
        self.root = Node()
        self.aabb_min = np.full(3, np.inf, np.float32)
        self.aabb_max = np.full(3, -np.inf, np.float32)
        self.centered = np.eye(4, dtype=np.float32)
        self.tri_count = 0
        self.vertex_count = 0


def _bbox_transform(bmin, bmax, m):
    cs = np.array([[x, y, z, 1.0] for x in (bmin[0], bmax[0]) for y in (bmin[1], bmax[1])
                   for z in (bmin[2], bmax[2])], np.float32)
    t = (m @ cs.T).T[:, :3]
    return t.min(axis=0), t.max(axis=0)


def _finish_model(model):
    """centeredTransform = scale(3/|diag|) * translate(-centre, y -> -min.y)  (ModelLoader.cpp:417-425)."""
    def walk(node, parent):
        cur = parent @ node.transform
        for mesh in node.meshes:
            lo, hi = _bbox_transform(mesh.aabb[0], mesh.aabb[1], cur)
            model.aabb_min = np.minimum(model.aabb_min, lo)
            model.aabb_max = np.maximum(model.aabb_max, hi)
            model.tri_count += len(mesh.indices) // 3
            model.vertex_count += len(mesh.vertices)
        for c in node.children:
            walk(c, cur)
    # the reference starts rootAABB at (0,0,0)-(0,0,0) (Geometry.h:30-31) and merges into it
    model.aabb_min = np.zeros(3, np.float32)
    model.aabb_max = np.zeros(3, np.float32)
    walk(model.root, np.eye(4, dtype=np.float32))
    trans = (model.aabb_max + model.aabb_min) / -2.0
    trans[1] = -model.aabb_min[1]
    ln = float(np.linalg.norm(model.aabb_max - model.aabb_min))
    s = np.eye(4, dtype=np.float32)
    s[0, 0] = s[1, 1] = s[2, 2] = 3.0 / ln
    t = np.eye(4, dtype=np.float32)
    t[:3, 3] = trans
    model.centered = (s @ t).astype(np.float32)
    return model


_WRAP = {10497: 0, 33648: 1, 33071: 2}  # REPEAT, MIRRORED_REPEAT, CLAMP_TO_EDGE


def load_gltf(path, image_cache=None):
    image_cache = {} if image_cache is None else image_cache
    base = os.path.dirname(path)
    with open(path) as f:
        g = json.load(f)
    buffers = []
    for b in g["buffers"]:
        with open(os.path.join(base, b["uri"]), "rb") as f:
            buffers.append(f.read())

    def accessor(idx):
        a = g["accessors"][idx]
        bv = g["bufferViews"][a["bufferView"]]
        dt = np.dtype(_COMP[a["componentType"]])
        nc = _NCOMP[a["type"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        buf = buffers[bv["buffer"]]
        cnt = a["count"]
        arr = np.lib.stride_tricks.as_strided(
            np.frombuffer(buf, dtype=dt, count=(stride * (cnt - 1)) // dt.itemsize + nc, offset=off),
            shape=(cnt, nc), strides=(stride, dt.itemsize))
        return np.array(arr)

    def image_for(tex_index):
        tex = g["textures"][tex_index]
        img = g["images"][tex["source"]]
        p = os.path.join(base, img["uri"])
        if p not in image_cache:
            image_cache[p] = load_image_rgba(p)
        samp = g.get("samplers", [{}])[tex["sampler"]] if "sampler" in tex and g.get("samplers") else {}
        return dict(path=p, image=image_cache[p], wrap_u=_WRAP.get(samp.get("wrapS", 10497), 0),
                    wrap_v=_WRAP.get(samp.get("wrapT", 10497), 0))

    def build_mesh(prim):
        at = prim["attributes"]
        pos = accessor(at["POSITION"]).astype(np.float32)
        nrm = accessor(at["NORMAL"]).astype(np.float32) if "NORMAL" in at else None
        uv = accessor(at["TEXCOORD_0"]).astype(np.float32) if "TEXCOORD_0" in at else np.zeros((len(pos), 2), np.float32)
        # assimp's glTF2 importer flips v once and aiProcess_FlipUVs flips it back: net effect is the file's own v
        uv = uv.copy()
        idx = accessor(prim["indices"]).astype(np.int32).reshape(-1) if "indices" in prim \
            else np.arange(len(pos), dtype=np.int32)
        if nrm is None:
            nrm = np.zeros_like(pos)
        if "TANGENT" in at:
            tan = accessor(at["TANGENT"]).astype(np.float32)[:, :3]
        else:
            tan = calc_tangents(pos, uv, nrm, idx)
        m = Mesh()
        m.vertices = make_vertices(pos, uv, nrm, tan)
        m.indices = idx
        m.aabb = (pos.min(axis=0), pos.max(axis=0))
        m.shading = "pbr"                               # glTF => Shading_PBR (ModelLoader.cpp:326-332)
        mat = g["materials"][prim["material"]] if "material" in prim else {}
        m.alpha_blend = mat.get("alphaMode", "OPAQUE") == "BLEND"
        m.double_sided = bool(mat.get("doubleSided", False))
        pbr = mat.get("pbrMetallicRoughness", {})
        if "baseColorTexture" in pbr:
            m.textures["albedo"] = image_for(pbr["baseColorTexture"]["index"])
        if "metallicRoughnessTexture" in pbr:
            m.textures["metal_roughness"] = image_for(pbr["metallicRoughnessTexture"]["index"])
        if "normalTexture" in mat:
            m.textures["normal"] = image_for(mat["normalTexture"]["index"])
        if "occlusionTexture" in mat:
            m.textures["ao"] = image_for(mat["occlusionTexture"]["index"])
        if "emissiveTexture" in mat:
            m.textures["emissive"] = image_for(mat["emissiveTexture"]["index"])
        return m

    def node_matrix(n):
        if "matrix" in n:
            return np.array(n["matrix"], np.float32).reshape(4, 4).T.copy()
        t = np.array(n.get("translation", [0, 0, 0]), np.float64)
        q = np.array(n.get("rotation", [0, 0, 0, 1]), np.float64)
        s = np.array(n.get("scale", [1, 1, 1]), np.float64)
        x, y, z, w = q
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        m = np.eye(4)
        m[:3, :3] = r * s[None, :]
        m[:3, 3] = t
        return m.astype(np.float32)

    def build_node(i):
        n = g["nodes"][i]
        node = Node()
        node.transform = node_matrix(n)
        if "mesh" in n:
            for prim in g["meshes"][n["mesh"]]["primitives"]:
                node.meshes.append(build_mesh(prim))
        for c in n.get("children", []):
            node.children.append(build_node(c))
        return node

    model = Model()
    roots = g["scenes"][g.get("scene", 0)]["nodes"]
    if len(roots) == 1:
        model.root = build_node(roots[0])
    else:
        for r in roots:
            model.root.children.append(build_node(r))
    return _finish_model(model)


def load_obj(path, image_cache=None):
    image_cache = {} if image_cache is None else image_cache
    base = os.path.dirname(path)
    vs, vts, vns, corners = [], [], [], []
    mtllib = None
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == "vt":
                vts.append([float(x) for x in p[1:3]])
            elif p[0] == "vn":
                vns.append([float(x) for x in p[1:4]])
            elif p[0] == "mtllib" and len(p) > 1:
                mtllib = p[1]
            elif p[0] == "f":
                face = []
                for c in p[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    ti = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    face.append((vi, ti, ni))
                for k in range(1, len(face) - 1):           # fan triangulation
                    corners.extend([face[0], face[k], face[k + 1]])
    vs = np.array(vs, np.float32)
    vts = np.array(vts, np.float32) if vts else np.zeros((1, 2), np.float32)
    vns = np.array(vns, np.float32) if vns else np.zeros((1, 3), np.float32)
    c = np.array(corners, np.int64)
    fix = lambda i, n: np.where(i > 0, i - 1, np.where(i < 0, n + i, 0))
    pos = vs[fix(c[:, 0], len(vs))]
    uv = vts[fix(c[:, 1], len(vts))].copy()
    uv[:, 1] = 1.0 - uv[:, 1]                             # aiProcess_FlipUVs
    nrm = vns[fix(c[:, 2], len(vns))]
    idx = np.arange(len(pos), dtype=np.int32)            # assimp OBJ: unshared corner vertices
    m = Mesh()
    m.vertices = make_vertices(pos, uv, nrm, calc_tangents(pos, uv, nrm, idx))
    m.indices = idx
    m.aabb = (pos.min(axis=0), pos.max(axis=0))
    m.shading = "blinnphong"                              # OBJ => Shading_BlinnPhong
    if mtllib and os.path.exists(os.path.join(base, mtllib)):
        with open(os.path.join(base, mtllib)) as f:
            for line in f:
                p = line.split()
                if len(p) >= 2 and p[0] in ("map_Kd", "map_Kn", "map_Bump", "norm"):
                    ip = os.path.join(base, p[-1])
                    if not os.path.exists(ip):
                        continue
                    if ip not in image_cache:
                        image_cache[ip] = load_image_rgba(ip)
                    key = "albedo" if p[0] == "map_Kd" else "normal"
                    m.textures[key] = dict(path=ip, image=image_cache[ip], wrap_u=0, wrap_v=0)
    model = Model()
    model.root.meshes.append(m)
    return _finish_model(model)


def load_model(path, image_cache=None):
    if path.lower().endswith(".obj"):
        return load_obj(path, image_cache)
    return load_gltf(path, image_cache)


def find_assets_dir():
    """Assets travel with the repo snapshot (assets/, git-ignored); never read /root/reference at run time."""
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.abspath(os.path.join(here, "..", ".."))
    for cand in (os.environ.get("SGL_ASSETS_DIR"), os.path.join(root, "assets")):
        if cand and os.path.isdir(cand) and os.path.exists(os.path.join(cand, "assets.json")):
            return cand
    return None
