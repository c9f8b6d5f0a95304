/* reference_types.h -- the handful of ChaNGa declarations that gravity.h and Ewald.cpp need from headers which
 * pull in Charm++ (TreeNode.h, GenericTreeNode.h, GravityParticle.h, ParallelGravity.h), given here with the
 * reference's names and meanings so that those two files compile UNMODIFIED.  MultipoleMoments.h itself is the
 * reference's own (its <pup.h>, <Vector3D.h>, <OrientedBox.h> resolve to the stand-ins in this directory).
 *
 * TEST INFRASTRUCTURE ONLY (oracle/gravity_ref.cpp, oracle/ewald_ref.cpp).  Field names as in GravityParticle.h
 * and GenericTreeNode.h; node types as GenericTreeNode.h:39-51; opening_geometry_factor as
 * TreeNode.h:35; EWT as ParallelGravity.h:636-639.  cosmoType is double (cosmoType.h without COSMO_FLOAT), the
 * scalar code path is the one compiled (CMK_SSE = 0, a build without --enable-sse2). */
#ifndef CB200_ORACLE_SHIM_REFERENCE_TYPES_H
#define CB200_ORACLE_SHIM_REFERENCE_TYPES_H

#include <cmath>
#include <cstddef>
using std::sqrt;

#include "cosmoType.h" /* the reference's */
#include "moments.h"   /* the reference's: FMOMR, MOMC, momEvalFmomrcm, momRescaleFmomr, momFmomr2Momc */
#include "Space.h"     /* oracle/shim/gravity */

namespace TreeStuff {
const double opening_geometry_factor = 2 / sqrt(3.0);
}

#define __SSEDEFS_H__ /* skips the reference's SSEdefs.h (include guard); the scalar path needs only CMK_SSE */
#ifndef CMK_SSE
#define CMK_SSE 0
#endif
#include "MultipoleMoments.h" /* the reference's, unmodified: the class, operator+= (parallel axis), the radius rules */

class ExternalGravityParticle {
 public:
  cosmoType mass;
  cosmoType soft;
  Vector3D<cosmoType> position;
};

class GravityParticle : public ExternalGravityParticle {
 public:
  Vector3D<cosmoType> treeAcceleration;
  cosmoType potential;
  cosmoType dtGrav;
  double interMass;
  int rung;
};

namespace Tree {
enum NodeType { Invalid = 1, Bucket, Internal, Boundary, NonLocal, Empty, Top, NonLocalBucket, Cached, CachedBucket, CachedEmpty };
class GenericTreeNode {
 public:
  NodeType type;
  MultipoleMoments moments;
  OrientedBox<cosmoType> boundingBox;
  int firstParticle, lastParticle;
  unsigned int particleCount;
  NodeType getType() const { return type; }
};
}  // namespace Tree

typedef struct ewaldTable {
  double hx, hy, hz;
  double hCfac, hSfac;
} EWT;

/* a cell record of the oracle (27 doubles, gravity_oracle.c CM_*: radius, soft, mass, cm[3], then
 * xx xy xz yy yz | xxx xyy xxy yyy xxz yyz xyz | xxxx xyyy xxxy yyyy xxxz yyyz xxyy xxyz xyyz) as a tree node */
inline void cb200_fill_node(Tree::GenericTreeNode &n, const double *c, const double *lo, const double *hi, int isBucket,
                            int first, int last, unsigned count) {
  n.type = isBucket ? Tree::Bucket : Tree::Internal;
  n.moments.setRadius(c[0]); n.moments.soft = c[1]; n.moments.totalMass = c[2];
  n.moments.cm = Vector3D<cosmoType>(c[3], c[4], c[5]);
  FMOMR &m = n.moments.mom;
  m.m = c[2];
  m.xx = c[6]; m.xy = c[7]; m.xz = c[8]; m.yy = c[9]; m.yz = c[10];
  m.xxx = c[11]; m.xyy = c[12]; m.xxy = c[13]; m.yyy = c[14]; m.xxz = c[15]; m.yyz = c[16]; m.xyz = c[17];
  m.xxxx = c[18]; m.xyyy = c[19]; m.xxxy = c[20]; m.yyyy = c[21]; m.xxxz = c[22]; m.yyyz = c[23];
  m.xxyy = c[24]; m.xxyz = c[25]; m.xyyz = c[26];
  if (lo && hi) {
    n.boundingBox.lesser_corner = Vector3D<cosmoType>(lo[0], lo[1], lo[2]);
    n.boundingBox.greater_corner = Vector3D<cosmoType>(hi[0], hi[1], hi[2]);
  }
  n.firstParticle = first; n.lastParticle = last; n.particleCount = count;
}

#endif
