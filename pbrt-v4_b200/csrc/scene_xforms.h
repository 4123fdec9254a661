// scene_xforms.h -- how the engine turns the transforms of the C ABI into rendering-space transforms (host only).
//
// Reference: rendering space is "camera-world" (CameraTransform, cameras.cpp:31-71: worldFromRender = Translate(camera position));
// BasicSceneBuilder hands every entity its RenderFromObject() = renderFromWorld * CTM (scene.h:475-477) and every instance
// RenderFromObject() * worldFromRender (scene.cpp:375-393).  A caller that sits in front of BasicSceneBuilder (this repo's .pbrt front
// end) passes world-space CTMs (BPT_SPACE_WORLD) and the engine composes them here exactly as the builder does; a caller that sits
// behind it (a binding that walks BasicScene, oracle/b200_binding.cpp) already holds the composed transforms and passes them as they
// are (BPT_SPACE_RENDER).  Shared by engine.cu and by the ABI recorder of the tests (tests/tools/abi_record.cpp), so that both
// routes can be compared matrix for matrix without a GPU.
#pragma once
#include <cstring>
#include "../../include/b200pt.h"
#include "xform.h"

namespace bpt {

inline M4 ToM4(const bpt_mat4 &a) { M4 r; std::memcpy(r.m, a.m, sizeof(r.m)); return r; }
inline bool AllZero(const bpt_mat4 &a) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (a.m[i][j] != 0) return false; return true; }
// A CTM as the reference holds it: the matrix with its analytically tracked inverse (Translate / Scale / Rotate / LookAt compose
// their exact inverses, util/transform.h:148-151, transform.cpp:36-117), so that products such as RenderFromObject() carry the
// reference's mInv bit for bit.  Callers that have no tracked inverse pass zeros and get the numerical one.
inline Xform TrackedXform(const bpt_mat4 &m, const bpt_mat4 &inv) { return AllZero(inv) ? Xform(ToM4(m)) : Xform(ToM4(m), ToM4(inv)); }

struct CameraSpaces { Xform cameraFromWorld, worldFromCamera, worldFromRender, renderFromWorld, renderFromCamera; };

// BasicSceneBuilder::Camera (scene.cpp:139-156) + CameraTransform (cameras.cpp:31-71, RenderingCoordinateSystem::CameraWorld)
inline bool ResolveCameraSpaces(const bpt_camera &cam, CameraSpaces *cs) {
    M4 cfw = ToM4(cam.camera_from_world), wfc;
    if (AllZero(cam.world_from_camera)) { if (!Inverse(cfw, &wfc)) return false; }
    else wfc = ToM4(cam.world_from_camera);
    cs->cameraFromWorld = Xform(cfw, wfc);
    cs->worldFromCamera = Inverse(cs->cameraFromWorld);
    Pt3 pCamera = cs->worldFromCamera(Pt3(0, 0, 0));
    cs->worldFromRender = Translate(Vec3(pCamera.x, pCamera.y, pCamera.z));
    cs->renderFromWorld = Inverse(cs->worldFromRender);
    cs->renderFromCamera = cs->renderFromWorld * cs->worldFromCamera;
    return true;
}
// shapes, lights, media: BasicSceneBuilder::RenderFromObject(index) (scene.h:475-477) =
//     Transform((renderFromWorld * graphicsState.ctm[index]).GetMatrix())
// -- a Transform rebuilt from the product MATRIX alone, so its inverse is the numerical Inverse() of that product
// (util/transform.h:44-56), not the product of the tracked inverses.  (Round 1 composed the tracked inverses: exact for pure
// translations and off in the last bits for anything with a rotation or a scale -- normals, instance rays and image-light
// directions go through mInv, which is what kept instanced and environment-lit scenes from being sample-exact, profiles/r02c.)
inline Xform ResolveRenderFromObject(int space, const CameraSpaces &cs, const bpt_mat4 &m, const bpt_mat4 &inv) {
    if (space == BPT_SPACE_RENDER) return TrackedXform(m, inv);
    return Xform(Mul(cs.renderFromWorld.m, ToM4(m)));
}
// instances: RenderFromObject(0) * worldFromRender with worldFromRender = Inverse(renderFromWorld) (scene.cpp:375-393)
inline Xform ResolveRenderFromInstance(int space, const CameraSpaces &cs, const bpt_mat4 &m, const bpt_mat4 &inv) {
    if (space == BPT_SPACE_RENDER) return TrackedXform(m, inv);
    return ResolveRenderFromObject(space, cs, m, inv) * Inverse(cs.renderFromWorld);
}

}  // namespace bpt
