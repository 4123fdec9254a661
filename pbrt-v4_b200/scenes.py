"""Seeded procedural scene generators for the BASELINE.json configs: each writes a
`.pbrt` file (plus binary PLY meshes) that both this engine and the reference `pbrt` binary read unchanged."""
import os
import struct

import numpy as np


def write_ply(path, P, idx, N=None):
    """binary_little_endian PLY with float x y z [nx ny nz] and int triangle faces."""
    P = np.asarray(P, dtype="<f4").reshape(-1, 3)
    idx = np.asarray(idx, dtype="<i4").reshape(-1, 3)
    with open(path, "wb") as f:
        hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n" % len(P)
        if N is not None:
            hdr += "property float nx\nproperty float ny\nproperty float nz\n"
        hdr += "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(idx)
        f.write(hdr.encode())
        if N is not None:
            f.write(np.hstack([P, np.asarray(N, dtype="<f4").reshape(-1, 3)]).astype("<f4").tobytes())
        else:
            f.write(P.tobytes())
        rec = np.empty(len(idx), dtype=[("n", "u1"), ("v", "<i4", 3)])
        rec["n"] = 3
        rec["v"] = idx
        f.write(rec.tobytes())


def _header(res, spp, maxdepth, cam, fov, extra=""):
    (ex, ey, ez), (lx, ly, lz) = cam
    return (f"LookAt {ex} {ey} {ez}  {lx} {ly} {lz}  0 1 0\n"
            f"Camera \"perspective\" \"float fov\" {fov}\n"
            f"Sampler \"zsobol\" \"integer pixelsamples\" {spp}\n"
            f"Integrator \"volpath\" \"integer maxdepth\" {maxdepth}\n"
            f"Film \"rgb\" \"integer xresolution\" {res[0]} \"integer yresolution\" {res[1]} \"string filename\" \"out.pfm\" \"bool savefp16\" false\n"
            f"{extra}WorldBegin\n")


def cornell(outdir, res=(256, 256), spp=64, maxdepth=5):
    """Config 1: Cornell box, 5 wall quads + 1 quad area light (12 triangles)."""
    os.makedirs(outdir, exist_ok=True)
    s = _header(res, spp, maxdepth, ((0, 1, 3.4), (0, 1, 0)), 40)
    quad = "Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [%s]\n"
    s += "AttributeBegin\n  AreaLightSource \"diffuse\" \"rgb L\" [17 12 4]\n  Material \"diffuse\" \"rgb reflectance\" [0 0 0]\n  "
    s += quad % "-0.25 1.98 -0.25  0.25 1.98 -0.25  0.25 1.98 0.25  -0.25 1.98 0.25"
    s += "AttributeEnd\nMaterial \"diffuse\" \"rgb reflectance\" [0.73 0.73 0.73]\n"
    s += quad % "-1 0 -1  -1 0 1  1 0 1  1 0 -1"
    s += quad % "-1 2 -1  1 2 -1  1 2 1  -1 2 1"
    s += quad % "-1 0 -1  1 0 -1  1 2 -1  -1 2 -1"
    s += "Material \"diffuse\" \"rgb reflectance\" [0.65 0.05 0.05]\n" + quad % "-1 0 -1  -1 2 -1  -1 2 1  -1 0 1"
    s += "Material \"diffuse\" \"rgb reflectance\" [0.12 0.45 0.15]\n" + quad % "1 0 -1  1 0 1  1 2 1  1 2 -1"
    path = os.path.join(outdir, "cornell.pbrt")
    open(path, "w").write(s)
    return path


def cornell_alpha(outdir, res=(96, 96), spp=16, maxdepth=5):
    """Cornell box with two alpha-tested panels (constant "float alpha" 0.5 and 0.25: GeometricPrimitive's stochastic alpha test,
    cpu/primitive.cpp:56-75), one of them inside an object instance so that the test hashes the instance-space ray."""
    path = cornell(outdir, res=res, spp=spp, maxdepth=maxdepth)
    s = open(path).read()
    s += ("Material \"diffuse\" \"rgb reflectance\" [0.2 0.3 0.8]\n"
          "Shape \"trianglemesh\" \"float alpha\" 0.5 \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [-0.6 0.3 0.2  0.2 0.3 0.2  0.2 1.3 0.2  -0.6 1.3 0.2]\n"
          "AttributeBegin\n  Material \"conductor\" \"float roughness\" 0.2\n  ObjectBegin \"panel\"\n"
          "    Shape \"trianglemesh\" \"float alpha\" 0.25 \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [-0.3 -0.3 0  0.3 -0.3 0  0.3 0.3 0  -0.3 0.3 0]\n"
          "  ObjectEnd\nAttributeEnd\n"
          "AttributeBegin\n  Translate 0.45 0.8 -0.2\n  Rotate 35 0.2 1 0.1\n  Scale 1.2 1.5 1\n  ObjectInstance \"panel\"\nAttributeEnd\n")
    open(path, "w").write(s)
    return path


LIGHT_SETS = {
    # point / spot / distant lights (lights.cpp:192-213, 1433-1464, 246-276), alone (a one-leaf light BVH), together with the quad
    # light (a real light BVH: bounds, cones, importance), and under transforms / with "power" / "illuminance" normalisation
    "point": 'LightSource "point" "point3 from" [0.3 1.6 0.4] "rgb I" [3 2.5 2]\n',
    "spot": 'LightSource "spot" "point3 from" [-0.5 1.9 0.6] "point3 to" [0.2 0 -0.3] "float coneangle" 35 "float conedeltaangle" 10 "rgb I" [12 12 10]\n',
    "distant": 'LightSource "distant" "point3 from" [0.4 1 0.7] "point3 to" [0 0 0] "rgb L" [1.5 1.4 1.2]\n',
    "all": ('AttributeBegin\n  Translate 0.1 0.2 0.05\n  Rotate 20 0 1 0\n  LightSource "point" "point3 from" [0.3 1.4 0.4] "blackbody I" 3200 "float power" 40\nAttributeEnd\n'
            'LightSource "spot" "point3 from" [-0.5 1.9 0.6] "point3 to" [0.2 0 -0.3] "float power" 60 "float scale" 0.5\n'
            'AttributeBegin\n  Rotate -15 1 0 0\n  LightSource "distant" "point3 to" [0.1 -1 -0.4] "float illuminance" 2\nAttributeEnd\n'
            'LightSource "point" "point3 from" [0 0.9 0.9] "float scale" 2\n'),
}


def cornell_lights(outdir, lights="all", area=True, res=(64, 64), spp=16, maxdepth=5):
    """Cornell box lit by LIGHT_SETS[lights] (in front of the open side the distant light and the off-centre point lights reach the
    walls directly), with or without the quad area light, plus a rough conductor panel that casts shadows."""
    path = cornell(outdir, res=res, spp=spp, maxdepth=maxdepth)
    s = open(path).read()
    if not area:
        a = s.index("AttributeBegin\n  AreaLightSource"); b = s.index("AttributeEnd\n", a) + len("AttributeEnd\n")
        s = s[:a] + s[b:]
    a = s.index("WorldBegin\n") + len("WorldBegin\n")
    s = s[:a] + LIGHT_SETS[lights] + s[a:]
    s += ("Material \"conductor\" \"float roughness\" 0.3\n"
          "Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [-0.5 0.55 -0.3  0.3 0.6 -0.45  0.35 0.62 0.25  -0.45 0.5 0.3]\n")
    open(path, "w").write(s)
    return path


QUADRIC_SETS = {
    # Sphere / Disk / Cylinder (shapes.h:107-790) in the Cornell box.  Each entry: (text inserted after WorldBegin, text appended, keep the quad light)
    "spheres": ("", (
        'AttributeBegin\n  Translate -0.45 0.35 -0.2\n  Material "diffuse" "rgb reflectance" [0.2 0.4 0.8]\n  Shape "sphere" "float radius" 0.35\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.45 0.4 0.3\n  Material "dielectric" "float eta" 1.5\n  Shape "sphere" "float radius" 0.3\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.1 1.2 -0.4\n  Rotate 40 1 0.3 0.2\n  Scale 1 0.7 1.2\n  Material "conductor" "float roughness" 0.15\n'
        '  Shape "sphere" "float radius" 0.3 "float zmin" -0.2 "float zmax" 0.25 "float phimax" 300\nAttributeEnd\n'), True),
    "disk_cylinder": ("", (
        'AttributeBegin\n  Translate -0.4 0 0.1\n  Rotate -90 1 0 0\n  Material "coateddiffuse" "rgb reflectance" [0.7 0.3 0.2] "float roughness" 0.05\n'
        '  Shape "cylinder" "float radius" 0.25 "float zmin" 0 "float zmax" 0.8\n  Shape "disk" "float radius" 0.25 "float height" 0.8\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.45 0.5 -0.1\n  Rotate 30 0 1 0\n  Rotate 70 1 0 0\n  Scale 1 1.3 -1\n  Material "conductor" "float roughness" 0.05\n'
        '  Shape "cylinder" "float radius" 0.2 "float zmin" -0.3 "float zmax" 0.4 "float phimax" 250\n'
        '  ReverseOrientation\n  Shape "disk" "float radius" 0.3 "float innerradius" 0.12 "float height" -0.3 "float phimax" 250\nAttributeEnd\n'), True),
    "lights": ("", (
        'AttributeBegin\n  Translate -0.5 1.5 0\n  AreaLightSource "diffuse" "rgb L" [9 7 4]\n  Shape "sphere" "float radius" 0.12\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.4 1.9 0.2\n  Rotate 100 1 0 0.2\n  AreaLightSource "diffuse" "rgb L" [6 8 10] "bool twosided" true\n  Shape "disk" "float radius" 0.25 "float innerradius" 0.05\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.5 0.3 0.3\n  Rotate 90 0 1 0\n  AreaLightSource "diffuse" "blackbody L" 3000 "float power" 30\n  Shape "cylinder" "float radius" 0.05 "float zmin" -0.4 "float zmax" 0.4\nAttributeEnd\n'
        'AttributeBegin\n  Translate -0.2 0.4 -0.3\n  Material "diffuse" "rgb reflectance" [0.6 0.6 0.6]\n  Shape "sphere" "float radius" 0.4\nAttributeEnd\n'
        # a large emissive sphere that contains part of the scene: reference points inside it take the area-sampling branch
        'AttributeBegin\n  Translate 0.6 0.2 -0.7\n  AreaLightSource "diffuse" "rgb L" [0.3 0.3 0.5] "bool twosided" true\n  ReverseOrientation\n  Shape "sphere" "float radius" 0.5 "float zmax" 0.3\nAttributeEnd\n'), False),
    "instanced": ("", (
        'AttributeBegin\n  Material "coatedconductor"\n  ObjectBegin "thing"\n    Shape "sphere" "float radius" 0.12\n'
        '    AttributeBegin\n      Translate 0 0 0.12\n      Shape "cylinder" "float radius" 0.05 "float zmin" 0 "float zmax" 0.25\n      Shape "disk" "float radius" 0.05 "float height" 0.25\n    AttributeEnd\n'
        '    Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [-0.2 -0.2 -0.12  0.2 -0.2 -0.12  0.2 0.2 -0.12  -0.2 0.2 -0.12]\n'
        '  ObjectEnd\nAttributeEnd\n'
        'AttributeBegin\n  Translate -0.45 0.5 0\n  Rotate 25 1 0 0.3\n  ObjectInstance "thing"\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.4 0.9 -0.2\n  Rotate -70 0.2 1 0\n  Scale 1.5 1.2 2\n  ObjectInstance "thing"\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.1 0.3 0.4\n  Scale -1 1 1\n  Rotate 200 0 0 1\n  ObjectInstance "thing"\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.5 0.25 0.5\n  Material "diffuse" "rgb reflectance" [0.3 0.6 0.3]\n  Shape "sphere" "float radius" 0.2\nAttributeEnd\n'), True),
    "media": ('MakeNamedMedium "fog" "string type" "homogeneous" "rgb sigma_a" [1 1.5 2] "rgb sigma_s" [6 6 5] "float g" 0.4 "float scale" 1.5\n', (
        'AttributeBegin\n  MediumInterface "fog" ""\n  Material "interface"\n  Translate -0.3 0.6 0\n  Shape "sphere" "float radius" 0.45\nAttributeEnd\n'
        'AttributeBegin\n  MediumInterface "fog" ""\n  Material "dielectric" "float eta" 1.33\n  Translate 0.5 0.35 0.2\n  Rotate -90 1 0 0\n'
        '  Shape "cylinder" "float radius" 0.25 "float zmin" -0.3 "float zmax" 0.3\n  Shape "disk" "float radius" 0.25 "float height" 0.3\n'
        '  ReverseOrientation\n  Shape "disk" "float radius" 0.25 "float height" -0.3\nAttributeEnd\n'), True),
    "mappings": ('AttributeBegin\n  Translate -0.4 0.45 -0.1\n  Rotate 30 0.3 1 0\n  Texture "sph" "spectrum" "imagemap" "string filename" "albedo.png" "string mapping" "spherical" "string filter" "trilinear"\nAttributeEnd\n'
                 'AttributeBegin\n  Translate 0.5 0 0.2\n  Rotate -90 1 0 0\n  Scale 1 1 0.5\n  Texture "cyl" "spectrum" "imagemap" "string filename" "colour.pfm" "string mapping" "cylindrical" "string filter" "ewa"\nAttributeEnd\n'
                 'AttributeBegin\n  Rotate 20 0 1 0\n  Texture "pla" "spectrum" "checkerboard" "string mapping" "planar" "vector3 v1" [3 0 0.5] "vector3 v2" [0 2 1] "float udelta" 0.25 "rgb tex1" [0.7 0.7 0.2] "rgb tex2" [0.1 0.2 0.6]\nAttributeEnd\n', (
        'AttributeBegin\n  Translate -0.4 0.45 -0.1\n  Rotate 30 0.3 1 0\n  Material "diffuse" "texture reflectance" "sph"\n  Shape "sphere" "float radius" 0.42\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.5 0 0.2\n  Rotate -90 1 0 0\n  Material "coateddiffuse" "texture reflectance" "cyl" "float roughness" 0.1\n'
        '  Shape "cylinder" "float radius" 0.22 "float zmin" 0 "float zmax" 0.9\n  Shape "disk" "float radius" 0.22 "float height" 0.9\nAttributeEnd\n'
        'Material "diffuse" "texture reflectance" "pla"\n'
        'Shape "bilinearmesh" "point3 P" [-0.2 0.02 0.3  0.45 0.02 0.5  -0.3 0.02 0.9  0.4 0.3 0.95]\n'
        'Material "conductor" "float roughness" 0\n'
        'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [0.95 0.2 -0.6  0.95 0.2 0.5  0.6 1.4 0.5  0.6 1.4 -0.6]\n'), True),
    "textured": ('Texture "alb" "spectrum" "imagemap" "string filename" "albedo.png" "float uscale" 4 "float vscale" 2\n'
                 'Texture "hm" "float" "imagemap" "string filename" "height.pfm" "float scale" 0.03 "float uscale" 3 "float vscale" 3\n', (
        'AttributeBegin\n  Translate -0.4 0.45 -0.1\n  Rotate 30 0.3 1 0\n  Material "diffuse" "texture reflectance" "alb" "texture displacement" "hm"\n  Shape "sphere" "float radius" 0.42\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.5 0 0.2\n  Rotate -90 1 0 0\n  Material "coateddiffuse" "texture reflectance" "alb" "float roughness" 0.1\n'
        '  Shape "cylinder" "float radius" 0.22 "float zmin" 0 "float zmax" 0.9\n  Shape "disk" "float radius" 0.22 "float height" 0.9\nAttributeEnd\n'
        'Material "conductor" "float roughness" 0\n'
        'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [0.95 0.2 -0.6  0.95 0.2 0.5  0.6 1.4 0.5  0.6 1.4 -0.6]\n'), True),
}


def cornell_quadrics(outdir, which="spheres", res=(64, 64), spp=16, maxdepth=5):
    """Cornell box with QUADRIC_SETS[which]: spheres / disks / cylinders as surfaces (diffuse, glass, conductors, clipped by zmin /
    zmax / phimax, under rotations, non-uniform and mirroring scales), as area lights, inside object instances, as the boundary of
    a participating medium, and carrying image textures and bump maps (their dndu / dndv are not zero)."""
    path = cornell(outdir, res=res, spp=spp, maxdepth=maxdepth)
    head, tail, keep_quad = QUADRIC_SETS[which]
    s = open(path).read()
    if which in ("textured", "mappings"):
        texture_images(outdir)
    if not keep_quad:
        a = s.index("AttributeBegin\n  AreaLightSource"); b = s.index("AttributeEnd\n", a) + len("AttributeEnd\n")
        s = s[:a] + s[b:]
    a = s.index("WorldBegin\n") + len("WorldBegin\n")
    s = s[:a] + head + s[a:] + tail
    open(path, "w").write(s)
    return path


def write_ply_faces(path, P, faces, N=None, UV=None):
    """binary_little_endian PLY whose faces are triangles and quads mixed (lists of 3 or 4 vertex indices); optional normals and uv."""
    P = np.asarray(P, dtype="<f4").reshape(-1, 3)
    cols = [P]
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n" % len(P)
    if N is not None:
        hdr += "property float nx\nproperty float ny\nproperty float nz\n"; cols.append(np.asarray(N, dtype="<f4").reshape(-1, 3))
    if UV is not None:
        hdr += "property float u\nproperty float v\n"; cols.append(np.asarray(UV, dtype="<f4").reshape(-1, 2))
    hdr += "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(faces)
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(np.hstack(cols).astype("<f4").tobytes())
        for face in faces:
            f.write(bytes([len(face)]) + np.asarray(face, dtype="<i4").tobytes())


BILINEAR_SETS = {
    # BilinearPatch (shapes.h:1350-1545) in the Cornell box: (text appended, keep the triangle quad light)
    "patches": ((
        # the ceiling light as a rectangular patch: spherical-rectangle sampling (shapes.cpp:1262-1300)
        'AttributeBegin\n  AreaLightSource "diffuse" "rgb L" [17 12 4]\n  Material "diffuse" "rgb reflectance" [0 0 0]\n'
        '  Shape "bilinearmesh" "point3 P" [-0.25 1.98 -0.25  -0.25 1.98 0.25  0.25 1.98 -0.25  0.25 1.98 0.25]\nAttributeEnd\n'
        # a twisted emissive patch (area sampling with the bilinear warp), two-sided
        'AttributeBegin\n  AreaLightSource "diffuse" "rgb L" [2 3 5] "bool twosided" true\n'
        '  Shape "bilinearmesh" "point3 P" [-0.9 0.1 -0.8  -0.5 0.15 -0.9  -0.95 0.7 -0.7  -0.55 0.6 -0.4]\nAttributeEnd\n'
        # a saddle with uv coordinates and vertex normals, under a mirroring transform
        'AttributeBegin\n  Translate 0.35 0.3 0.1\n  Rotate 25 0 1 0\n  Scale -1 1 1\n  Material "coateddiffuse" "rgb reflectance" [0.6 0.3 0.2] "float roughness" 0.1\n'
        '  Shape "bilinearmesh" "integer indices" [0 1 3 4  1 2 4 5] "point3 P" [-0.4 0 -0.3  0 0.25 -0.3  0.4 0 -0.3  -0.4 0.3 0.3  0 0 0.3  0.4 0.35 0.3]\n'
        '    "point2 uv" [0 0  0.5 0  1 0.1  0 1  0.5 0.9  1 1] "normal N" [0.2 1 0.1  0 1 0.3  -0.2 1 0.1  0.1 1 -0.2  0 1 -0.1  -0.3 1 0]\nAttributeEnd\n'
        # a patch collapsed to a triangle and a glass sheet
        'AttributeBegin\n  Material "dielectric" "float eta" 1.5\n  ReverseOrientation\n'
        '  Shape "bilinearmesh" "point3 P" [-0.6 0.05 0.5  -0.1 0.05 0.6  -0.6 0.9 0.4  -0.1 0.9 0.6]\nAttributeEnd\n'
        'AttributeBegin\n  Material "conductor" "float roughness" 0.2\n  Shape "bilinearmesh" "point3 P" [0.5 1.2 -0.6  0.9 1.2 -0.6  0.7 1.7 -0.7  0.9 1.2 -0.6]\nAttributeEnd\n'), False),
    "ply": ((
        'AttributeBegin\n  Translate -0.35 0.45 -0.1\n  Rotate 30 1 1 0\n  Material "diffuse" "rgb reflectance" [0.3 0.5 0.7]\n  Shape "plymesh" "string filename" "quads.ply"\nAttributeEnd\n'
        'AttributeBegin\n  Material "conductor" "float roughness" 0.1\n  ObjectBegin "q"\n    Shape "plymesh" "string filename" "quads.ply"\n  ObjectEnd\nAttributeEnd\n'
        'AttributeBegin\n  Translate 0.45 0.4 0.2\n  Rotate -50 0 1 0.2\n  Scale 0.8 1.1 0.8\n  ObjectInstance "q"\nAttributeEnd\n'), True),
}


def cornell_bilinear(outdir, which="patches", res=(64, 64), spp=16, maxdepth=5):
    """Cornell box with BILINEAR_SETS[which]: bilinear patches as surfaces and as area lights (rectangles sampled as spherical
    rectangles, other patches by area), with uv / N, collapsed to triangles, and the quad faces of a PLY file (a box whose sides are
    quads and whose caps are triangle fans: triangle mesh + bilinear patch mesh over the same vertices), top-level and instanced."""
    path = cornell(outdir, res=res, spp=spp, maxdepth=maxdepth)
    tail, keep_quad = BILINEAR_SETS[which]
    s = open(path).read()
    if which == "ply":
        # an open prism: 8 ring vertices bottom / top + 2 cap centres; sides = quads, caps = triangles; smooth-ish normals, uv
        k = 8
        ang = np.arange(k) * 2 * np.pi / k
        ring = np.stack([0.25 * np.cos(ang), 0.25 * np.sin(ang)], 1)
        P = [[x, -0.3, z] for x, z in ring] + [[1.3 * x, 0.3, 1.3 * z] for x, z in ring] + [[0, -0.3, 0], [0, 0.3, 0]]
        N = [[x, -0.1, z] for x, z in ring] + [[x, 0.2, z] for x, z in ring] + [[0, -1, 0], [0, 1, 0]]
        UV = [[i / k, 0] for i in range(k)] + [[i / k, 1] for i in range(k)] + [[0.5, 0], [0.5, 1]]
        faces = []
        for i in range(k):
            j = (i + 1) % k
            faces.append([i, j, k + j, k + i])
            faces.append([2 * k, j, i])
            faces.append([2 * k + 1, k + i, k + j])
        write_ply_faces(os.path.join(outdir, "quads.ply"), P, faces, N=N, UV=UV)
    if not keep_quad:
        a = s.index("AttributeBegin\n  AreaLightSource"); b = s.index("AttributeEnd\n", a) + len("AttributeEnd\n")
        s = s[:a] + s[b:]
    s += tail
    open(path, "w").write(s)
    return path


def write_png(path, img):
    """8-bit PNG (grey HxW, grey+alpha HxWx2, RGB HxWx3 or RGBA HxWx4), filter type 0 rows, zlib level 6."""
    import struct, zlib
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = b"".join(b"\0" + img[y].tobytes() for y in range(h))
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def write_pfm(path, img):
    """PFM ("Pf" one channel HxW, "PF" RGB HxWx3), little-endian, bottom row first."""
    img = np.ascontiguousarray(img, dtype="<f4")
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(("%s\n%d %d\n-1.0\n" % ("Pf" if img.ndim == 2 else "PF", w, h)).encode())
        f.write(img[::-1].tobytes())


def texture_images(outdir, seed=5):
    """The images of the textured test scenes: an RGB PNG whose sides are not powers of two (the pyramid is built from a resized
    copy), a grey PNG, an RGBA PNG with an opaque alpha channel, a one-channel float PFM height map and an RGB float PFM."""
    rng = np.random.Generator(np.random.PCG64(seed))
    os.makedirs(outdir, exist_ok=True)
    y, x = np.mgrid[0:37, 0:53]
    checker = (((x // 6) + (y // 5)) % 2).astype(np.float64)
    rgb = np.stack([0.15 + 0.7 * checker, 0.25 + 0.5 * np.sin(x / 5.0) ** 2, 0.2 + 0.6 * (y / 36.0)], -1) + 0.08 * rng.random((37, 53, 3))
    write_png(os.path.join(outdir, "albedo.png"), np.clip(rgb * 255, 0, 255))
    g = 0.5 + 0.5 * np.sin(np.mgrid[0:32, 0:32][1] / 3.0) * np.cos(np.mgrid[0:32, 0:32][0] / 4.0)
    write_png(os.path.join(outdir, "grey.png"), np.clip(g * 255, 0, 255))
    rgba = np.concatenate([np.clip(rng.random((16, 16, 3)) * 255, 0, 255), np.full((16, 16, 1), 255.0)], -1)
    write_png(os.path.join(outdir, "rgba.png"), rgba)
    yy, xx = np.mgrid[0:64, 0:64] / 64.0
    height = 0.03 * (np.sin(14 * xx) * np.sin(11 * yy) + 0.3 * rng.random((64, 64)))
    write_pfm(os.path.join(outdir, "height.pfm"), height)
    write_pfm(os.path.join(outdir, "colour.pfm"), 0.1 + 0.8 * rng.random((24, 40, 3)))
    ny, nx = np.mgrid[0:20, 0:28]
    nrm = np.stack([0.35 * np.sin(nx / 2.5), 0.35 * np.cos(ny / 2.0), np.ones((20, 28))], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    write_png(os.path.join(outdir, "normals.png"), np.clip((0.5 + 0.5 * nrm) * 255, 0, 255))


TEXTURE_SETS = {
    # what changes on the Cornell box: (Texture directives, the floor / back-wall material, extra geometry)
    "albedo": ('Texture "alb" "spectrum" "imagemap" "string filename" "albedo.png"\n', 'Material "diffuse" "texture reflectance" "alb"\n'),
    "trilinear": ('Texture "alb" "spectrum" "imagemap" "string filename" "albedo.png" "string filter" "trilinear" "float uscale" 3 "float vscale" 2 "float udelta" 0.25\n',
                  'Material "diffuse" "texture reflectance" "alb"\n'),
    "ewa": ('Texture "alb" "spectrum" "imagemap" "string filename" "colour.pfm" "string filter" "ewa" "float uscale" 4 "float vscale" 4 "string wrap" "clamp" "float scale" 0.9\n',
            'Material "diffuse" "texture reflectance" "alb"\n'),
    "point": ('Texture "alb" "spectrum" "imagemap" "string filename" "rgba.png" "string filter" "point" "float uscale" 2 "float vscale" 2 "bool invert" true "string encoding" "linear"\n',
              'Material "coateddiffuse" "texture reflectance" "alb" "float roughness" 0.2\n'),
    "grey": ('Texture "alb" "spectrum" "imagemap" "string filename" "grey.png" "string wrap" "black" "float uscale" 1.5 "float vscale" 1.5\n',
             'Material "diffusetransmission" "texture reflectance" "alb" "texture transmittance" "alb" "float scale" 0.8\n'),
    "roughness": ('Texture "rough" "float" "imagemap" "string filename" "grey.png" "float scale" 0.4\nTexture "tint" "spectrum" "imagemap" "string filename" "colour.pfm"\n',
                  'Material "conductor" "texture roughness" "rough" "texture reflectance" "tint"\n'),
    "bump": ('Texture "h" "float" "imagemap" "string filename" "height.pfm" "float uscale" 2 "float vscale" 2\nTexture "alb" "spectrum" "imagemap" "string filename" "albedo.png"\n',
             'Material "diffuse" "texture reflectance" "alb" "texture displacement" "h"\n'),
    "scale_mix": ('Texture "alb" "spectrum" "imagemap" "string filename" "albedo.png"\nTexture "m" "float" "imagemap" "string filename" "grey.png" "float uscale" 2\n'
                  'Texture "half" "spectrum" "scale" "texture tex" "alb" "float scale" 0.6\nTexture "tinted" "spectrum" "scale" "texture tex" "alb" "texture scale" "m"\n'
                  'Texture "both" "spectrum" "mix" "texture tex1" "half" "rgb tex2" [0.1 0.6 0.2] "texture amount" "m"\nTexture "one" "spectrum" "scale" "texture tex" "both" "float scale" 1\n'
                  'Texture "r" "float" "scale" "texture tex" "m" "float scale" 0.5\nTexture "r2" "float" "mix" "texture tex1" "r" "float tex2" 0.05 "float amount" 0.3\n'
                  'Texture "c" "spectrum" "constant" "rgb value" [0.7 0.6 0.3]\n',
                  'Material "coateddiffuse" "texture reflectance" "one" "texture roughness" "r2"\n'),
    "checkerboard": ('Texture "alb" "spectrum" "imagemap" "string filename" "albedo.png"\n'
                     'Texture "chk" "spectrum" "checkerboard" "float uscale" 7 "float vscale" 5.5 "float udelta" 0.1 "texture tex1" "alb" "rgb tex2" [0.08 0.1 0.3]\n'
                     'Texture "fchk" "float" "checkerboard" "float uscale" 3 "float vscale" 3 "float tex1" 0.02 "float tex2" 0.35\n',
                     'Material "conductor" "texture reflectance" "chk" "texture roughness" "fchk"\n'),
    "normalmap": ('Texture "alb" "spectrum" "imagemap" "string filename" "albedo.png"\n', 'Material "diffuse" "texture reflectance" "alb" "string normalmap" "normals.png"\n'),
    "bump_dielectric": ('Texture "h" "float" "imagemap" "string filename" "height.pfm" "float uscale" 3 "float vscale" 3 "string filter" "trilinear"\n',
                        'Material "dielectric" "float eta" 1.5 "float roughness" 0.15 "texture displacement" "h"\n'),
}


def cornell_textured(outdir, which="albedo", res=(64, 64), spp=16, maxdepth=5, mirror=True):
    """Cornell box whose floor and back wall carry TEXTURE_SETS[which] (the meshes get uv coordinates and, for the back wall, vertex
    normals, so that dndu / dndv are not zero), seen directly and -- with mirror -- through a smooth conductor panel and a glass
    slab, which is where the specular ray differentials of SurfaceInteraction::SpawnRay end up in a texture lookup."""
    path = cornell(outdir, res=res, spp=spp, maxdepth=maxdepth)
    texture_images(outdir)
    s = open(path).read()
    texdef, mat = TEXTURE_SETS[which]
    a = s.index("WorldBegin\n") + len("WorldBegin\n")
    s = s[:a] + texdef + s[a:]
    floor = 'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [-1 0 -1  -1 0 1  1 0 1  1 0 -1]\n'
    back = 'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [-1 0 -1  1 0 -1  1 2 -1  -1 2 -1]\n'
    assert floor in s and back in s
    s = s.replace(floor, "AttributeBegin\n" + mat + floor[:-1] + ' "point2 uv" [0 0  0 1  1 1  1 0]\nAttributeEnd\n')
    s = s.replace(back, "AttributeBegin\n" + mat + back[:-1] + ' "point2 uv" [0 0  1 0  1 1  0 1] "normal N" [-0.2 0.1 1  0.2 0.1 1  0.2 -0.1 1  -0.2 -0.1 1]\nAttributeEnd\n')
    if mirror:
        s += ('Material "conductor" "float roughness" 0\n'
              'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [-0.95 0.2 -0.6  -0.95 0.2 0.5  -0.6 1.4 0.5  -0.6 1.4 -0.6]\n'
              'Material "dielectric" "float eta" 1.45\n'
              'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3  4 6 5 4 7 6  0 4 5 0 5 1  1 5 6 1 6 2  2 6 7 2 7 3  3 7 4 3 4 0]\n'
              '  "point3 P" [0.2 0.02 0.1  0.8 0.02 0.1  0.8 0.02 0.7  0.2 0.02 0.7   0.2 0.5 0.1  0.8 0.5 0.1  0.8 0.5 0.7  0.2 0.5 0.7]\n')
    open(path, "w").write(s)
    return path


def heightfield_mesh(n, seed=1234, amp=0.05, noise=0.01):
    g = np.linspace(-1.0, 1.0, n, dtype=np.float64)
    X, Z = np.meshgrid(g, g, indexing="xy")
    rng = np.random.Generator(np.random.PCG64(seed))
    Y = amp * np.sin(9 * X) * np.cos(7 * Z) + noise * rng.standard_normal(X.shape)
    P = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = np.arange(n - 1)
    I, J = np.meshgrid(i, i, indexing="xy")
    v00 = (J * n + I).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + n
    v11 = v01 + 1
    # counter-clockwise seen from +y so that geometric normals point up
    idx = np.concatenate([np.stack([v00, v01, v11], -1), np.stack([v00, v11, v10], -1)], axis=0).astype(np.int32)
    return P, idx


def diffuse_heightfield(outdir, n=708, res=(1024, 1024), spp=256, maxdepth=5, seed=1234):
    """Config 2: (n-1)^2*2-triangle diffuse height field (n=708 -> 999 698 triangles), one 1x1 quad light at y=2."""
    os.makedirs(outdir, exist_ok=True)
    P, idx = heightfield_mesh(n, seed)
    write_ply(os.path.join(outdir, "heightfield.ply"), P, idx)
    s = _header(res, spp, maxdepth, ((0, 1.6, 2.6), (0, 0, 0)), 45)
    s += ("AttributeBegin\n  AreaLightSource \"diffuse\" \"rgb L\" [20 20 20]\n  Material \"diffuse\" \"rgb reflectance\" [0 0 0]\n"
          "  Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [-0.5 2 -0.5  0.5 2 -0.5  0.5 2 0.5  -0.5 2 0.5]\n"
          "AttributeEnd\n")
    s += "Material \"diffuse\" \"rgb reflectance\" [0.6 0.6 0.6]\nShape \"plymesh\" \"string filename\" \"heightfield.ply\"\n"
    path = os.path.join(outdir, "heightfield.pbrt")
    open(path, "w").write(s)
    return path


def write_envmap_pfm(path, res=256, seed=11):
    """Procedural HDR environment map in pbrt's equal-area octahedral parameterisation: sky gradient + 3 Gaussian 'suns'."""
    from .imageio import write_pfm
    g = (np.arange(res, dtype=np.float64) + 0.5) / res
    U, V = np.meshgrid(g, g, indexing="xy")
    # equal-area square -> sphere (Clarberg), vectorised; only used to author the picture (the renderers do their own mapping)
    u, v = 2 * U - 1, 2 * V - 1
    up, vp = np.abs(u), np.abs(v)
    sd = 1 - (up + vp)
    r = 1 - np.abs(sd)
    phi = np.where(r == 0, 1.0, (vp - up) / np.where(r == 0, 1, r) + 1) * np.pi / 4
    z = np.copysign(1 - r * r, sd)
    c, s = np.copysign(np.cos(phi), u), np.copysign(np.sin(phi), v)
    x, y = c * r * np.sqrt(np.maximum(0, 2 - r * r)), s * r * np.sqrt(np.maximum(0, 2 - r * r))
    img = np.zeros((res, res, 3))
    t = 0.5 * (z + 1)
    img += (0.15 + 0.85 * t)[..., None] * np.array([0.35, 0.55, 1.0])
    rng = np.random.Generator(np.random.PCG64(seed))
    for k in range(3):
        d = rng.normal(size=3); d[2] = abs(d[2]) + 0.3; d /= np.linalg.norm(d)
        cosang = x * d[0] + y * d[1] + z * d[2]
        img += (5e4 if k == 0 else 5e2) * np.exp((cosang - 1) / (2e-4 if k == 0 else 4e-3))[..., None] * np.array([1.0, 0.9, 0.7])
    write_pfm(path, img.astype(np.float32))


def mixed_envmap(outdir, n=2236, res=(1920, 1080), spp=1024, maxdepth=5, envres=2048, nmesh=64, seed=1234):
    """Config 3: height field split into `nmesh` strips alternating dielectric (eta 1.5, smooth) / conductor (Cu, roughness
    0.05 and 0.3), lit by a procedural HDR environment map."""
    os.makedirs(outdir, exist_ok=True)
    P, idx = heightfield_mesh(n, seed)
    write_envmap_pfm(os.path.join(outdir, "env.pfm"), envres)
    s = _header(res, spp, maxdepth, ((0, 1.6, 2.6), (0, 0, 0)), 45)
    s += "AttributeBegin\n  Rotate -90 1 0 0\n  LightSource \"infinite\" \"string filename\" \"env.pfm\"\nAttributeEnd\n"
    tri_per = (len(idx) + nmesh - 1) // nmesh
    mats = ["Material \"dielectric\" \"float eta\" 1.5 \"float roughness\" 0\n",
            "Material \"conductor\" \"spectrum eta\" \"metal-Cu-eta\" \"spectrum k\" \"metal-Cu-k\" \"float roughness\" 0.05\n",
            "Material \"dielectric\" \"float eta\" 1.5 \"float roughness\" 0\n",
            "Material \"conductor\" \"spectrum eta\" \"metal-Cu-eta\" \"spectrum k\" \"metal-Cu-k\" \"float roughness\" 0.3\n"]
    for m in range(nmesh):
        sub = idx[m * tri_per:(m + 1) * tri_per]
        if len(sub) == 0:
            continue
        used, inv = np.unique(sub.reshape(-1), return_inverse=True)
        write_ply(os.path.join(outdir, f"strip{m:03d}.ply"), P[used], inv.reshape(-1, 3))
        s += mats[m % 4] + f"Shape \"plymesh\" \"string filename\" \"strip{m:03d}.ply\"\n"
    # a diffuse ground plane below so that refracted paths terminate on something
    s += ("Material \"diffuse\" \"rgb reflectance\" [0.4 0.4 0.4]\n"
          "Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [-3 -0.5 -3  -3 -0.5 3  3 -0.5 3  3 -0.5 -3]\n")
    path = os.path.join(outdir, "mixed.pbrt")
    open(path, "w").write(s)
    return path


def icosphere(subdiv):
    t = (1.0 + 5 ** 0.5) / 2
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    V = [np.array(v, dtype=np.float64) / np.linalg.norm(v) for v in V]
    for _ in range(subdiv):
        cache, F2 = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = V[a] + V[b]
                V.append(m / np.linalg.norm(m))
                cache[key] = len(V) - 1
            return cache[key]
        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            F2 += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        F = F2
    return np.array(V, dtype=np.float32), np.array(F, dtype=np.int32)


def instanced(outdir, n_inst=100000, subdiv=3, res=(2048, 2048), spp=512, maxdepth=5, seed=99, material="coatedconductor", parts="trs"):
    """Config 4: one icosphere mesh (20*4^subdiv triangles; subdiv=3 -> 1280) as an object definition, `n_inst` instances with
    seeded random translate / rotate / scale in a 100 x 10 x 100 slab (scaled with n_inst), one quad light above."""
    os.makedirs(outdir, exist_ok=True)
    V, F = icosphere(subdiv)
    write_ply(os.path.join(outdir, "ico.ply"), V, F)
    ext = 50.0 * (n_inst / 100000.0) ** (1 / 3)
    s = _header(res, spp, maxdepth, ((0, 0.9 * ext, 1.6 * ext), (0, 0, 0)), 45)
    L = 40.0
    s += ("AttributeBegin\n  AreaLightSource \"diffuse\" \"rgb L\" [%g %g %g]\n  Material \"diffuse\" \"rgb reflectance\" [0 0 0]\n"
          "  Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [%g %g %g  %g %g %g  %g %g %g  %g %g %g]\nAttributeEnd\n" %
          (L, L, L, -ext, ext, -ext, ext, ext, -ext, ext, ext, ext, -ext, ext, ext))
    s += ("Material \"diffuse\" \"rgb reflectance\" [0.5 0.5 0.5]\nShape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" "
          "[%g %g %g  %g %g %g  %g %g %g  %g %g %g]\n" % (-2 * ext, -0.12 * ext, -2 * ext, -2 * ext, -0.12 * ext, 2 * ext, 2 * ext, -0.12 * ext, 2 * ext, 2 * ext, -0.12 * ext, -2 * ext))
    mat = {"coatedconductor": "Material \"coatedconductor\"\n", "diffuse": "Material \"diffuse\" \"rgb reflectance\" [0.7 0.4 0.3]\n",
           "conductor": "Material \"conductor\" \"float roughness\" 0.2\n"}[material]
    s += "AttributeBegin\n" + mat + "ObjectBegin \"ico\"\n  Shape \"plymesh\" \"string filename\" \"ico.ply\"\nObjectEnd\nAttributeEnd\n"
    rng = np.random.Generator(np.random.PCG64(seed))
    lines = []
    for i in range(n_inst):
        tx, ty, tz = rng.uniform(-ext, ext), rng.uniform(-0.1 * ext, 0.1 * ext), rng.uniform(-ext, ext)
        ang = rng.uniform(0, 360); ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        sc = rng.uniform(0.35, 0.7) * (ext / 50.0) * 50.0 / (n_inst ** (1 / 3)) * 0.9
        # `parts` selects the directives of the instance transform (parity bisection: translation / rotation / scale alone)
        lines.append("AttributeBegin\n" + ("  Translate %.6f %.6f %.6f\n" % (tx, ty, tz) if "t" in parts else "") +
                     ("  Rotate %.5f %.6f %.6f %.6f\n" % (ang, ax[0], ax[1], ax[2]) if "r" in parts else "") +
                     ("  Scale %.6f %.6f %.6f\n" % (sc, sc, sc) if "s" in parts else "") + "  ObjectInstance \"ico\"\nAttributeEnd\n")
    s += "".join(lines)
    path = os.path.join(outdir, "instanced.pbrt")
    open(path, "w").write(s)
    return path


def smoke_density(grid=256, seed=7):
    """fBm-thresholded sphere: value noise in 5 octaves (seeded), kept inside a sphere of radius 0.45 around the centre."""
    rng = np.random.default_rng(seed)
    d = np.zeros((grid, grid, grid), np.float32)
    amp, tot = 1.0, 0.0
    for octave in range(5):
        n = min(grid, 4 << octave)
        coarse = rng.random((n + 1, n + 1, n + 1), dtype=np.float32)
        # trilinear upsampling of the coarse lattice to the grid
        t = (np.arange(grid, dtype=np.float32) + 0.5) / grid * n
        i0 = np.minimum(t.astype(np.int32), n - 1); f = (t - i0).astype(np.float32)
        def ax(a, axis):
            lo = np.take(a, i0, axis=axis); hi = np.take(a, i0 + 1, axis=axis)
            shape = [1, 1, 1]; shape[axis] = grid
            w = f.reshape(shape)
            return lo * (1 - w) + hi * w
        c = ax(ax(ax(coarse, 0), 1), 2)
        d += amp * c; tot += amp; amp *= 0.5
    d /= tot
    z, y, x = np.meshgrid(*[(np.arange(grid, dtype=np.float32) + 0.5) / grid - 0.5] * 3, indexing="ij")
    r = np.sqrt(x * x + y * y + z * z)
    d = np.clip((d - 0.42) * 6.0, 0, 1) * np.clip((0.45 - r) * 12.0, 0, 1)
    return d.astype(np.float32)   # indexed [z][y][x]: x fastest, as GridMedium expects


def smoke(outdir, grid=256, res=(1024, 1024), spp=2048, maxdepth=5, seed=7, homogeneous=False, emission=None):
    """Config 5: grid-density smoke (sigma_a = 2, sigma_s = 8, g = 0.3) inside a unit box with the "interface" material,
    diffuse floor, quad light; null-scattering VolPath.  `emission`: None, "le" (an RGB "Le" on the medium; homogeneous media also get
    a "Lescale") or "grid" (uniformgrid: "Le" plus a per-voxel "Lescale" array = the density, so that the dense core glows)."""
    os.makedirs(outdir, exist_ok=True)
    s = _header(res, spp, maxdepth, ((0, 1.1, 3.0), (0, 0.55, 0)), 38)
    quad = "Shape \"trianglemesh\" \"integer indices\" [0 1 2 0 2 3] \"point3 P\" [%s]\n"
    s += "AttributeBegin\n  Translate -0.5 0.05 -0.5\n"
    if homogeneous:
        s += "  MakeNamedMedium \"smoke\" \"string type\" \"homogeneous\" \"rgb sigma_a\" [2 2 2] \"rgb sigma_s\" [8 8 8] \"float g\" 0.3 \"float scale\" 0.25\n"
        if emission:
            s += "    \"rgb Le\" [1.0 0.55 0.2] \"float Lescale\" 3\n"
    else:
        d = smoke_density(grid, seed)
        s += (f"  MakeNamedMedium \"smoke\" \"string type\" \"uniformgrid\" \"rgb sigma_a\" [2 2 2] \"rgb sigma_s\" [8 8 8] \"float g\" 0.3\n"
              f"    \"integer nx\" {grid} \"integer ny\" {grid} \"integer nz\" {grid} \"point3 p0\" [0 0 0] \"point3 p1\" [1 1 1]\n    \"float density\" [\n")
        flat = d.reshape(-1)
        s += "\n".join(" ".join(repr(float(v)) for v in flat[i:i + 16]) for i in range(0, flat.size, 16))
        s += "\n    ]\n"
        if emission:
            s += "    \"rgb Le\" [1.0 0.55 0.2]\n"
        if emission == "grid":
            s += "    \"float Lescale\" [\n" + "\n".join(" ".join(repr(float(4 * v)) for v in flat[i:i + 16]) for i in range(0, flat.size, 16)) + "\n    ]\n"
    s += "AttributeEnd\n"
    s += "AttributeBegin\n  AreaLightSource \"diffuse\" \"rgb L\" [30 28 24]\n  Material \"diffuse\" \"rgb reflectance\" [0 0 0]\n  "
    s += quad % "-0.4 2.2 -0.4  0.4 2.2 -0.4  0.4 2.2 0.4  -0.4 2.2 0.4"
    s += "AttributeEnd\nMaterial \"diffuse\" \"rgb reflectance\" [0.6 0.6 0.6]\n"
    s += quad % "-3 0 -3  -3 0 3  3 0 3  3 0 -3"
    # unit box with outward-facing triangles, "interface" material, smoke inside / vacuum outside
    s += "AttributeBegin\n  MediumInterface \"smoke\" \"\"\n  Material \"interface\"\n  Translate -0.5 0.05 -0.5\n"
    s += ("  Shape \"trianglemesh\" \"point3 P\" [0 0 0  1 0 0  1 1 0  0 1 0  0 0 1  1 0 1  1 1 1  0 1 1]\n"
          "    \"integer indices\" [0 2 1 0 3 2  4 5 6 4 6 7  0 1 5 0 5 4  3 6 2 3 7 6  0 4 7 0 7 3  1 2 6 1 6 5]\n")
    s += "AttributeEnd\n"
    path = os.path.join(outdir, "smoke.pbrt")
    open(path, "w").write(s)
    return path


def smoke_rgb(outdir, grid=12, res=(64, 64), spp=16, maxdepth=5, seed=7, emission=True):
    """The smoke scene with an "rgbgrid" medium (RGBGridMedium, media.h:355-427): coloured absorption and scattering per voxel and,
    with emission, an "Le" grid -- the same box, floor and light as smoke()."""
    path = smoke(outdir, grid=grid, res=res, spp=spp, maxdepth=maxdepth, seed=seed)
    s = open(path).read()
    a = s.index("  MakeNamedMedium"); b = s.index("AttributeEnd\n", a)
    d = smoke_density(grid, seed).reshape(-1)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    tint = 0.5 + 0.5 * rng.random((d.size, 3))
    def arr(v):
        flat = np.asarray(v, dtype=np.float64).reshape(-1)
        return "\n".join(" ".join(repr(float(x)) for x in flat[i:i + 12]) for i in range(0, flat.size, 12))
    med = (f"  MakeNamedMedium \"smoke\" \"string type\" \"rgbgrid\" \"float g\" 0.3 \"float scale\" 1.5 \"integer nx\" {grid} \"integer ny\" {grid} \"integer nz\" {grid}\n"
           f"    \"point3 p0\" [0 0 0] \"point3 p1\" [1 1 1]\n    \"rgb sigma_a\" [\n{arr(2 * d[:, None] * tint)}\n    ]\n    \"rgb sigma_s\" [\n{arr(8 * d[:, None] * tint[:, ::-1])}\n    ]\n")
    if emission:
        med += f"    \"rgb Le\" [\n{arr(d[:, None] ** 2 * np.array([2.0, 0.9, 0.3]))}\n    ] \"float Lescale\" 1.5\n"
    s = s[:a] + med + s[b:]
    open(path, "w").write(s)
    return path


def as_gbuffer(path, coordinate_system="camera"):
    """Rewrites a generated scene to use Film "gbuffer" (output name *.exr); returns the new path."""
    s = open(path).read().replace('Film "rgb"', f'Film "gbuffer" "string coordinatesystem" "{coordinate_system}"').replace('"out.pfm"', '"out.exr"')
    out = os.path.join(os.path.dirname(path), "gbuffer_" + os.path.basename(path))
    open(out, "w").write(s)
    return out
