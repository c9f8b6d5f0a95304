/* Space.h -- stand-in for the header of that name in N-BodyShop/utility (structures/), an un-vendored submodule
 * that is absent from /root/reference (configure looks for it in ../utility/structures, Makefile.in:40).
 *
 * TEST INFRASTRUCTURE ONLY: it exists so that the reference's OWN gravity.h compiles here unmodified
 * (oracle/gravity_ref.cpp; likewise MultipoleMoments.h and Ewald.cpp).  It holds the few geometric types they touch --
 * Vector3D, Sphere, OrientedBox, with component-wise arithmetic -- and the three predicates of namespace Space that
 * gravity.h calls.  Two of them follow the reference's own restatement for
 * its GPU walk, which IS in the tree: cuda_intersect (box / sphere, CUDAMoments.cu:137-159) and CUDA_intersect
 * (sphere / sphere, CUDAMoments.cu:161-168).  The third, contained(box, sphere), has no copy in the tree (the
 * reference's GPU walk only has bucket targets, which never ask it): it is the farthest-corner test of the
 * library, restated from its published source -- the one piece of this pin that rests on a restatement. */
#ifndef CB200_ORACLE_SHIM_SPACE_H
#define CB200_ORACLE_SHIM_SPACE_H

#include <cmath>

template <typename T>
struct Vector3D {
  T x, y, z;
  Vector3D() : x(0), y(0), z(0) {}
  Vector3D(T a, T b, T c) : x(a), y(b), z(c) {}
  template <typename U>
  Vector3D(const Vector3D<U> &o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
  T lengthSquared() const { return x * x + y * y + z * z; }
  Vector3D operator+(const Vector3D &o) const { return Vector3D(x + o.x, y + o.y, z + o.z); }
  Vector3D operator-(const Vector3D &o) const { return Vector3D(x - o.x, y - o.y, z - o.z); }
  Vector3D operator-() const { return Vector3D(-x, -y, -z); }
  Vector3D operator*(T s) const { return Vector3D(x * s, y * s, z * s); }
  Vector3D operator/(T s) const { return Vector3D(x / s, y / s, z / s); }
  Vector3D &operator+=(const Vector3D &o) { x += o.x; y += o.y; z += o.z; return *this; }
  Vector3D &operator-=(const Vector3D &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  T length() const { return std::sqrt(lengthSquared()); }
};
template <typename T, typename S>
inline Vector3D<T> operator*(const S &s, const Vector3D<T> &v) {
  return Vector3D<T>(s * v.x, s * v.y, s * v.z);
}

template <typename T>
struct Sphere {
  Vector3D<T> origin;
  T radius;
  Sphere(const Vector3D<T> &o, T r) : origin(o), radius(r) {}
};

template <typename T>
struct OrientedBox {
  Vector3D<T> lesser_corner, greater_corner;
};

namespace Space {

/* the sphere reaches the box: squared distance from the centre to the box, axis by axis, with the early exits */
template <typename T>
inline bool intersect(const OrientedBox<T> &b, const Sphere<T> &s) {
  T dsq = 0;
  const T rsq = s.radius * s.radius;
  T delta;
  if ((delta = b.lesser_corner.x - s.origin.x) > 0) dsq += delta * delta;
  else if ((delta = s.origin.x - b.greater_corner.x) > 0) dsq += delta * delta;
  if (rsq < dsq) return false;
  if ((delta = b.lesser_corner.y - s.origin.y) > 0) dsq += delta * delta;
  else if ((delta = s.origin.y - b.greater_corner.y) > 0) dsq += delta * delta;
  if (rsq < dsq) return false;
  if ((delta = b.lesser_corner.z - s.origin.z) > 0) dsq += delta * delta;
  else if ((delta = s.origin.z - b.greater_corner.z) > 0) dsq += delta * delta;
  return dsq <= s.radius * s.radius;
}

/* two spheres touch */
template <typename T>
inline bool intersect(const Sphere<T> &a, const Sphere<T> &b) {
  const Vector3D<T> d = a.origin - b.origin;
  return d.lengthSquared() <= (a.radius + b.radius) * (a.radius + b.radius);
}

/* the whole box lies inside the sphere: its farthest corner does */
template <typename T>
inline bool contained(const OrientedBox<T> &b, const Sphere<T> &s) {
  T dsq = 0;
  T d1, d2;
  d1 = b.lesser_corner.x - s.origin.x; d2 = b.greater_corner.x - s.origin.x;
  dsq += (d1 * d1 > d2 * d2) ? d1 * d1 : d2 * d2;
  d1 = b.lesser_corner.y - s.origin.y; d2 = b.greater_corner.y - s.origin.y;
  dsq += (d1 * d1 > d2 * d2) ? d1 * d1 : d2 * d2;
  d1 = b.lesser_corner.z - s.origin.z; d2 = b.greater_corner.z - s.origin.z;
  dsq += (d1 * d1 > d2 * d2) ? d1 * d1 : d2 * d2;
  return dsq <= s.radius * s.radius;
}

}  // namespace Space
#endif
