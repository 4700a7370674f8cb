// intersectMesh / intersectScene (reference: src/raytracer/intersection.h:9-18), evaluated
// on the GPU through mcskin_cuda_intersect.
#pragma once

#include "math/ray.h"
#include "scene/mesh.h"
#include "scene/scene.h"
#include "scene/triangle.h"

HitResult intersectMesh(const Ray& ray, const Mesh& mesh);
HitResult intersectScene(const Ray& ray, const Scene& scene);
