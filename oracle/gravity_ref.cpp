/* gravity_ref.cpp -- the reference's OWN gravity.h, compiled here unmodified, behind a small C interface.
 *
 * TEST INFRASTRUCTURE ONLY (built by oracle/Makefile into oracle/_ref/libgravity_ref.so when /root/reference is
 * present; nothing is copied out of the reference tree).  It pins the oracle's restatements of
 *   SPLINE / SPLINEQ                  gravity.h:20-66, 147-182
 *   partBucketForce                   gravity.h:267-303
 *   nodeBucketForce (hexadecapole)    gravity.h:399-475   (calls the reference's moments.c, linked in)
 *   openSoftening                     gravity.h:251-260
 *   openCriterionBucket / ...Node     gravity.h:597-723
 *   the moment build                  MultipoleMoments.h (operator+=, the radius rules), in the order of
 *                                     GenericTreeNode.h:220-255 and TreePiece.cpp:3505-3542
 * to the code they restate (tests/test_oracle_pins.py; golden vectors in tests/golden/gravity_kat.npz carry the pin
 * to the GPU box).
 *
 * gravity.h includes four headers.  Three exist in the reference and pull in Charm++ (TreeNode.h,
 * GenericTreeNode.h, SSEdefs.h): their include guards are defined below, so the preprocessor skips them, and the
 * handful of declarations gravity.h needs from them are given in oracle/shim/gravity/reference_types.h with the
 * reference's names and meanings -- field names as in GravityParticle.h / GenericTreeNode.h, node types as
 * GenericTreeNode.h:39-51, opening_geometry_factor as TreeNode.h:35; MultipoleMoments is the reference's own class
 * (MultipoleMoments.h, unmodified).  The fourth, Space.h, belongs to the absent utility/structures submodule:
 * oracle/shim/gravity/Space.h.  The scalar (non-SSE) code path is the one compiled: CMK_SSE = 0, as in a build
 * without --enable-sse2; cosmoType is double (cosmoType.h without COSMO_FLOAT). */
#define TREENODE_H
#define GENERICTREENODE_H
#define __SSEDEFS_H__
#define CMK_SSE 0
#ifndef HEXADECAPOLE
#define HEXADECAPOLE 1
#endif

#include <cstring>
#include <vector>

#include "reference_types.h" /* oracle/shim/gravity: the declarations named above */

cosmoType theta = 0.7;
cosmoType thetaMono = 0.7 * 0.7 * 0.7 * 0.7;

#include "gravity.h" /* the reference's, unmodified */

/* ------------------------------------------------------------------------------------------------------------
 * C interface in the oracle's array layouts: a particle row is {mass, soft, x, y, z}, an accumulator row
 * {ax, ay, az, pot, dtGrav}, a cell record the 27 doubles of gravity_oracle.c (CM_*: radius, soft, mass, cm[3],
 * then xx xy xz yy yz | xxx xyy xxy yyy xxz yyz xyz | xxxx xyyy xxxy yyyy xxxz yyyz xxyy xxyz xyyz). */
namespace {

/* rows [first, last] of the oracle's arrays as the reference's particle array (indexed like the rows) */
std::vector<GravityParticle> load_particles(const double *part, const double *vars, const unsigned char *rung, int first,
                                            int last) {
  std::vector<GravityParticle> p((size_t)last + 1);
  for (int j = first; j <= last; ++j) {
    const double *r = part + (size_t)j * 5, *v = vars + (size_t)j * 5;
    p[j].mass = r[0]; p[j].soft = r[1]; p[j].position = Vector3D<cosmoType>(r[2], r[3], r[4]);
    p[j].treeAcceleration = Vector3D<cosmoType>(v[0], v[1], v[2]);
    p[j].potential = v[3]; p[j].dtGrav = v[4]; p[j].interMass = 0.0;
    p[j].rung = rung ? rung[j] : 0;
  }
  return p;
}
void store_particles(const std::vector<GravityParticle> &p, double *vars, int first, int last) {
  for (int j = first; j <= last; ++j) {
    double *v = vars + (size_t)j * 5;
    v[0] = p[j].treeAcceleration.x; v[1] = p[j].treeAcceleration.y; v[2] = p[j].treeAcceleration.z;
    v[3] = p[j].potential; v[4] = p[j].dtGrav;
  }
}

}  // namespace

extern "C" {

void gref_set_theta(double t, double tMono) { theta = t; thetaMono = tMono; }

void gref_spline(double r2, double twoh, double *a, double *b) {
  cosmoType aa, bb;
  SPLINE(r2, twoh, aa, bb);
  *a = aa; *b = bb;
}
void gref_splineq(double invr, double r2, double twoh, double *abcd) {
  cosmoType a, b, c, d;
  SPLINEQ(invr, r2, twoh, a, b, c, d);
  abcd[0] = a; abcd[1] = b; abcd[2] = c; abcd[3] = d;
}

/* one source particle {mass, soft, x, y, z} on the target rows [first, last]; returns the reference's count */
int gref_part_bucket_force(const double *src, const double *shift, const double *part, int first, int last,
                           const unsigned char *rung, int activeRung, double *vars) {
  ExternalGravityParticle s;
  s.mass = src[0]; s.soft = src[1]; s.position = Vector3D<cosmoType>(src[2], src[3], src[4]);
  Tree::GenericTreeNode req;
  req.type = Tree::Bucket; req.firstParticle = first; req.lastParticle = last; req.particleCount = (unsigned)(last - first + 1);
  std::vector<GravityParticle> p = load_particles(part, vars, rung, first, last);
  const int n = partBucketForce(&s, &req, p.data(), Vector3D<cosmoType>(shift[0], shift[1], shift[2]), activeRung);
  store_particles(p, vars, first, last);
  return n;
}

/* one source cell on the target bucket [first, last] with moments my27 and box [mylo, myhi]: the whole
 * nodeBucketForce, softened branch included (a cell whose softening reaches the bucket acts as a particle) */
int gref_node_bucket_force(const double *cell27, const double *shift, const double *my27, const double *mylo,
                           const double *myhi, const double *part, int first, int last, const unsigned char *rung,
                           int activeRung, double *vars) {
  Tree::GenericTreeNode node, req;
  cb200_fill_node(node, cell27, nullptr, nullptr, 0, 0, -1, 1000);
  cb200_fill_node(req, my27, mylo, myhi, 1, first, last, (unsigned)(last - first + 1));
  std::vector<GravityParticle> p = load_particles(part, vars, rung, first, last);
  const int n = nodeBucketForce(&node, &req, p.data(), Vector3D<cosmoType>(shift[0], shift[1], shift[2]), activeRung);
  store_particles(p, vars, first, last);
  return n;
}

int gref_open_softening(const double *node27, const double *shift, const double *my27, const double *mylo,
                        const double *myhi) {
  Tree::GenericTreeNode node, my;
  cb200_fill_node(node, node27, nullptr, nullptr, 0, 0, -1, 1000);
  cb200_fill_node(my, my27, mylo, myhi, 0, 0, -1, 1000);
  return openSoftening(&node, &my, Vector3D<cosmoType>(shift[0], shift[1], shift[2]));
}

int gref_open_criterion_node(const double *node27, int nodeParticleCount, const double *shift, const double *my27,
                             const double *mylo, const double *myhi, int myIsBucket) {
  Tree::GenericTreeNode node, my;
  cb200_fill_node(node, node27, nullptr, nullptr, 0, 0, -1, (unsigned)nodeParticleCount);
  cb200_fill_node(my, my27, mylo, myhi, myIsBucket, 0, -1, 1000);
  return openCriterionNode(&node, &my, Vector3D<cosmoType>(shift[0], shift[1], shift[2]));
}

int gref_open_criterion_bucket(const double *node27, int nodeParticleCount, const double *shift, const double *my27,
                               const double *mylo, const double *myhi) {
  Tree::GenericTreeNode node, my;
  cb200_fill_node(node, node27, nullptr, nullptr, 0, 0, -1, (unsigned)nodeParticleCount);
  cb200_fill_node(my, my27, mylo, myhi, 1, 0, -1, 1000);
  return openCriterionBucket(&node, &my, Vector3D<cosmoType>(shift[0], shift[1], shift[2])) ? 1 : 0;
}

/* The moment build with the reference's own MultipoleMoments (MultipoleMoments.h, unmodified): every operation is
 * the reference's -- operator+=(particle), operator+=(moments), calculateRadiusBox / FirstParticle / FarthestParticle
 * / FarthestCorner -- in the order the reference applies them: GenericTreeNode::makeBucket (GenericTreeNode.h:220-255)
 * for a bucket, accumulateMomentsFromChild for each child then calculateRadiusFarthestCorner on the tight box
 * (TreePiece.cpp:3505-3510, 3540-3542; 2966-3002 in the local build) for an internal node.  Same arguments as
 * orc_build_moments: children have larger indices than their parent; out = numNodes x 27 (CM_* order). */
void gref_build_moments(const double *pos, const double *mass, const double *soft, const int *child0, const int *child1,
                        const int *firstPart, const int *lastPart, const double *geolo, const double *geohi,
                        const double *boxlo, const double *boxhi, int numNodes, double *out) {
  std::vector<MultipoleMoments> all((size_t)numNodes);
  for (int i = numNodes - 1; i >= 0; --i) {
    MultipoleMoments &m = all[i];
    OrientedBox<double> geo, tight;
    geo.lesser_corner = Vector3D<double>(geolo[3 * i], geolo[3 * i + 1], geolo[3 * i + 2]);
    geo.greater_corner = Vector3D<double>(geohi[3 * i], geohi[3 * i + 1], geohi[3 * i + 2]);
    tight.lesser_corner = Vector3D<double>(boxlo[3 * i], boxlo[3 * i + 1], boxlo[3 * i + 2]);
    tight.greater_corner = Vector3D<double>(boxhi[3 * i], boxhi[3 * i + 1], boxhi[3 * i + 2]);
    if (child0[i] < 0 && child1[i] < 0) {
      const int first = firstPart[i], last = lastPart[i], count = last - first + 1;
      std::vector<GravityParticle> part((size_t)count);
      for (int j = 0; j < count; ++j) {
        const int k = first + j;
        part[j].mass = mass[k]; part[j].soft = soft[k];
        part[j].position = Vector3D<cosmoType>(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]);
      }
      calculateRadiusBox(m, geo);
      if (m.getRadius() <= 0.0) {
        if (count > 1) calculateRadiusFirstParticle(m, part.data(), part.data() + count);
        else m.setRadius(1.0);
      }
      for (int j = 0; j < count; ++j) m += part[j];
      if (count > 1) calculateRadiusFarthestParticle(m, part.data(), part.data() + count);
    } else {
      if (child0[i] >= 0) m += all[child0[i]];
      if (child1[i] >= 0) m += all[child1[i]];
      calculateRadiusFarthestCorner(m, tight);
    }
    double *c = out + (size_t)i * 27;
    const FMOMR &f = m.mom;
    c[0] = m.getRadius(); c[1] = m.soft; c[2] = m.totalMass; c[3] = m.cm.x; c[4] = m.cm.y; c[5] = m.cm.z;
    c[6] = f.xx; c[7] = f.xy; c[8] = f.xz; c[9] = f.yy; c[10] = f.yz;
    c[11] = f.xxx; c[12] = f.xyy; c[13] = f.xxy; c[14] = f.yyy; c[15] = f.xxz; c[16] = f.yyz; c[17] = f.xyz;
    c[18] = f.xxxx; c[19] = f.xyyy; c[20] = f.xxxy; c[21] = f.yyyy; c[22] = f.xxxz; c[23] = f.yyyz;
    c[24] = f.xxyy; c[25] = f.xxyz; c[26] = f.xyyz;
  }
}

} /* extern "C" */
