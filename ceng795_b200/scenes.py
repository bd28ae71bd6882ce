"""Seeded synthetic scene generators for BASELINE.json's configs.

The reference repository ships no scenes (`.gitignore:34-38` excludes them), so
both arms — the untouched reference binary and the CUDA core — render XML
written here, in each tree's own dialect.  Floats are printed with `%.9g`, which
round-trips binary32 exactly through either parser.
"""
from __future__ import annotations

import io

import numpy as np

SEED = 795


def _f(a):
    return " ".join("%.9g" % float(x) for x in np.asarray(a, dtype=np.float32).ravel())


def _rows(a, fmt):
    buf = io.StringIO()
    np.savetxt(buf, a, fmt=fmt)
    return buf.getvalue()


def _uv_sphere(center, radius, slices, stacks):
    """Closed UV sphere: 2*slices*(stacks-1) triangles, CCW seen from outside."""
    c = np.asarray(center, np.float32)
    verts = [c + np.array([0, radius, 0], np.float32)]
    for i in range(1, stacks):
        th = np.pi * i / stacks
        for j in range(slices):
            ph = 2 * np.pi * j / slices
            verts.append(c + radius * np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], np.float32))
    verts.append(c - np.array([0, radius, 0], np.float32))
    verts = np.asarray(verts, np.float32)
    faces = []
    ring = lambda i, j: 1 + (i - 1) * slices + (j % slices)
    for j in range(slices):
        faces.append((0, ring(1, j + 1), ring(1, j)))
    for i in range(1, stacks - 1):
        for j in range(slices):
            a, b, c2, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces.append((a, b, c2))
            faces.append((b, d, c2))
    last = len(verts) - 1
    for j in range(slices):
        faces.append((last, ring(stacks - 1, j), ring(stacks - 1, j + 1)))
    return verts, np.asarray(faces, np.int32)


def _material_xml(mid, amb, dif, spec, phong, mirror=None, transparency=None, refr=None, roughness=None):
    s = f'    <Material id="{mid}">\n'
    s += f"      <AmbientReflectance>{_f(amb)}</AmbientReflectance>\n"
    s += f"      <DiffuseReflectance>{_f(dif)}</DiffuseReflectance>\n"
    s += f"      <SpecularReflectance>{_f(spec)}</SpecularReflectance>\n"
    if mirror is not None:
        s += f"      <MirrorReflectance>{_f(mirror)}</MirrorReflectance>\n"
    if roughness is not None:
        s += f"      <Roughness>{_f([roughness])}</Roughness>\n"
    s += f"      <PhongExponent>{_f([phong])}</PhongExponent>\n"
    if transparency is not None:
        s += f"      <Transparency>{_f(transparency)}</Transparency>\n"
        s += f"      <RefractionIndex>{_f([refr])}</RefractionIndex>\n"
    s += "    </Material>\n"
    return s


def cfg1_hw1(width=800, height=800, n_spheres=64, sphere_mesh=(20, 26), depth=6, seed=SEED,
             image_name="cfg1.png") -> str:
    """Config 1 — HW1 dialect: spheres + 2 triangle meshes + stand-alone triangles,
    3 point lights, Blinn-Phong + mirror, 1 spp."""
    rng = np.random.default_rng(seed)
    verts = []
    nv = 0

    def push(v):
        nonlocal nv
        base = nv
        verts.append(np.asarray(v, np.float32))
        nv += len(v)
        return base

    mats = ""
    for m in range(6):
        dif = rng.uniform(0.2, 0.9, 3)
        spec = rng.uniform(0.1, 0.6, 3)
        mirror = np.full(3, 0.5) if m % 2 == 1 else None
        mats += _material_xml(m + 1, rng.uniform(0.05, 0.2, 3), dif, spec, float(rng.choice([8, 20, 50, 100])), mirror)
    # room: 10 stand-alone triangles (floor, back, left, right, ceiling), facing inwards
    R = 6.0
    room = np.array([[-R, -R, -R], [R, -R, -R], [R, -R, R], [-R, -R, R],
                     [-R, R, -R], [R, R, -R], [R, R, R], [-R, R, R]], np.float32)
    rb = push(room)
    quads = [(0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (4, 5, 6, 7)]
    tri_xml = ""
    tid = 1
    for q in quads:
        for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3])):
            idx = [rb + t + 1 for t in tri]
            tri_xml += (f'    <Triangle id="{tid}">\n      <Material>{1 + (tid % 6)}</Material>\n'
                        f"      <Indices>{idx[0]} {idx[1]} {idx[2]}</Indices>\n    </Triangle>\n")
            tid += 1
    mesh_xml = ""
    for k, c in enumerate([(-2.5, -3.0, -1.0), (2.5, -2.0, 0.5)]):
        v, f = _uv_sphere(c, 1.6 + 0.3 * k, sphere_mesh[0], sphere_mesh[1])
        b = push(v)
        mesh_xml += (f'    <Mesh id="{k + 1}">\n      <Material>{2 + 3 * k}</Material>\n      <Faces>\n'
                     + _rows(f + b + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    sph_xml = ""
    centers = rng.uniform(-4, 4, (n_spheres, 3)).astype(np.float32)
    radii = rng.uniform(0.2, 0.6, n_spheres).astype(np.float32)
    cb = push(centers)
    for i in range(n_spheres):
        sph_xml += (f'    <Sphere id="{i + 1}">\n      <Material>{1 + int(rng.integers(0, 6))}</Material>\n'
                    f"      <Center>{cb + i + 1}</Center>\n      <Radius>{_f([radii[i]])}</Radius>\n    </Sphere>\n")
    lights = ""
    for i in range(3):
        pos = np.array([rng.uniform(-4, 4), rng.uniform(3, 5.5), rng.uniform(-1, 5)])
        inten = rng.uniform(5e2, 2e3, 3)
        lights += (f'    <PointLight id="{i + 1}">\n      <Position>{_f(pos)}</Position>\n'
                   f"      <Intensity>{_f(inten)}</Intensity>\n    </PointLight>\n")
    allv = np.concatenate(verts, 0)
    aspect = width / height
    return ("<Scene>\n  <BackgroundColor>10 20 40</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 0.5 5.9</Position>\n      <Gaze>0 -0.05 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-aspect, aspect, -1, 1])}</NearPlane>\n      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <ImageName>{image_name}</ImageName>\n"
            "    </Camera>\n  </Cameras>\n  <Lights>\n    <AmbientLight>25 25 25</AmbientLight>\n" + lights +
            "  </Lights>\n  <Materials>\n" + mats + "  </Materials>\n  <VertexData>\n" + _rows(allv, "%.9g") +
            "  </VertexData>\n  <Objects>\n" + mesh_xml + tri_xml + sph_xml + "  </Objects>\n</Scene>\n")


def heightfield(n, seed=SEED, extent=4.0):
    """(n+1)^2 vertices, 2*n*n triangles: y = .25 sin3x cos3z + .02 U(-1,1)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-extent, extent, n + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = 0.25 * np.sin(3 * X) * np.cos(3 * Z) + 0.02 * rng.uniform(-1, 1, X.shape)
    v = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    a = (j * (n + 1) + i).ravel()
    b = a + 1
    c = a + (n + 1)
    d = c + 1
    # CCW seen from +y
    f = np.empty((2 * n * n, 3), np.int64)
    f[0::2] = np.stack([a, c, b], -1)
    f[1::2] = np.stack([b, c, d], -1)
    return v, f.astype(np.int32)


def cfg2_hw3(n=724, width=1920, height=1080, n_instances=8, depth=2, seed=SEED, image_name="cfg2.png",
             smooth=False, cam_angle_deg=0.0) -> str:
    """Config 2 — HW2-era features in the HW3 dialect (the HW2 tree lacks MeshInstance):
    a (2*n*n)-triangle heightfield (n=724 -> 1 048 352), `n_instances` transformed
    instances of it, 4 spheres, 2 point lights, mirror depth 2, 1 spp."""
    rng = np.random.default_rng(seed + 2)
    v, f = heightfield(n, seed)
    sph_c = np.array([[-2.0, 1.0, 1.5], [2.2, 0.9, 0.5], [0.0, 1.4, -1.5], [1.0, 0.7, 2.5]], np.float32)
    sph_r = [0.6, 0.5, 0.8, 0.35]
    allv = np.concatenate([v, sph_c], 0)
    cb = len(v)
    mats = (_material_xml(1, [0.3, 0.3, 0.3], [0.45, 0.55, 0.35], [0.2, 0.2, 0.2], 20)
            + _material_xml(2, [0.1, 0.1, 0.1], [0.15, 0.15, 0.2], [0.5, 0.5, 0.5], 80, mirror=[0.6, 0.6, 0.6])
            + _material_xml(3, [0.1, 0.1, 0.1], [0.3, 0.1, 0.1], [0.4, 0.4, 0.4], 50, mirror=[0.35, 0.3, 0.3])
            + _material_xml(4, [0.3, 0.3, 0.3], [0.7, 0.3, 0.2], [0.3, 0.3, 0.3], 10)
            + _material_xml(5, [0.3, 0.3, 0.3], [0.2, 0.4, 0.8], [0.6, 0.6, 0.6], 40))
    xf = "  <Transformations>\n"
    inst = ""
    for k in range(n_instances):
        s = rng.uniform(0.08, 0.18)
        xf += f'    <Scaling id="{k + 1}">{_f([s, s * rng.uniform(1.0, 3.0), s])}</Scaling>\n'
        ang = np.linspace(0, 2 * np.pi, n_instances, endpoint=False)[k]
        xf += (f'    <Translation id="{k + 1}">{_f([2.6 * np.cos(ang), rng.uniform(1.0, 2.4), 2.6 * np.sin(ang) - 0.3])}'
               "</Translation>\n")
        axis = rng.normal(size=3)
        xf += f'    <Rotation id="{k + 1}">{_f([rng.uniform(0, 360), *axis])}</Rotation>\n'
        inst += (f'    <MeshInstance id="{k + 1}" baseMeshId="1" resetTransform="true">\n'
                 f"      <Material>{2 + (k % 4)}</Material>\n"
                 f"      <Transformations>s{k + 1} r{k + 1} t{k + 1}</Transformations>\n    </MeshInstance>\n")
    xf += "  </Transformations>\n"
    sph = ""
    for i in range(4):
        sph += (f'    <Sphere id="{i + 1}">\n      <Material>{[2, 4, 3, 5][i]}</Material>\n'
                f"      <Center>{cb + i + 1}</Center>\n      <Radius>{_f([sph_r[i]])}</Radius>\n    </Sphere>\n")
    aspect = width / height
    shading = ' shadingMode="smooth"' if smooth else ""
    ca, sa = np.cos(np.radians(cam_angle_deg)), np.sin(np.radians(cam_angle_deg))
    cam_pos = [5.6 * sa, 3.1, 5.6 * ca]       # orbit about +y (angle 0 = the canonical camera)
    cam_gaze = [-sa, -0.62, -ca]
    return ("<Scene>\n  <BackgroundColor>30 40 60</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            f"      <Position>{_f(cam_pos)}</Position>\n      <Gaze>{_f(cam_gaze)}</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.45 * aspect, 0.45 * aspect, -0.45, 0.45])}</NearPlane>\n"
            "      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <NumSamples>1</NumSamples>\n"
            f"      <ImageName>{image_name}</ImageName>\n    </Camera>\n  </Cameras>\n"
            "  <Lights>\n    <AmbientLight>25 25 25</AmbientLight>\n"
            '    <PointLight id="1">\n      <Position>-3 6 4</Position>\n      <Intensity>5000 4800 4500</Intensity>\n    </PointLight>\n'
            '    <PointLight id="2">\n      <Position>4 5 -2</Position>\n      <Intensity>2500 3000 3500</Intensity>\n    </PointLight>\n'
            "  </Lights>\n  <Materials>\n" + mats + "  </Materials>\n" + xf + "  <VertexData>\n"
            + _rows(allv, "%.9g") + "  </VertexData>\n  <Objects>\n"
            + f'    <Mesh id="1"{shading}>\n      <Material>1</Material>\n      <Faces>\n' + _rows(f + 1, "%d")
            + "      </Faces>\n    </Mesh>\n" + inst + sph + "  </Objects>\n</Scene>\n")
