/* device_layout.cuh -- how tree cells and particles live in HBM.
 *
 * The records ChaNGa hands us (CudaMultipoleMoments 27 reals, CompactPartData
 * 5 reals; include/changa_b200_types.h) are neither 16-byte aligned nor a
 * multiple of 16 bytes, so a warp cannot fetch them with 128-bit accesses.
 * DataManagerTransferLocalTree / ...RemoteChunk therefore upload the caller's
 * AoS records into a scratch buffer and a repack kernel rewrites them into
 * the two layouts below; the pointers handed back to the caller (and later
 * found again in CudaRequest::d_localMoments etc.) address the packed arrays.
 *
 *   PackedCell  32 reals = 8 (float) or 16 (double) 16-byte pieces, one
 *               128/256-byte row per cell, row-aligned: a cell is exactly one
 *               L2 line (float).  Multipole components are pre-multiplied by
 *               the (2l-1)!! factors of g2,g3,g4 (3, 15, 105) and the three
 *               trace combinations the evaluation needs are precomputed, so
 *               the inner loop is pure FMA work.
 *   PackedPart   8 reals: {x, y, z, mass | soft, 0, 0, 0}: one 32-byte sector
 *               (float).  p-c targets read the first half only.
 */
#ifndef CB200_DEVICE_LAYOUT_CUH
#define CB200_DEVICE_LAYOUT_CUH

#include "../../include/changa_b200_types.h"

namespace cb200 {

typedef cudatype real;

/* slots of a PackedCell */
enum {
  PK_CX, PK_CY, PK_CZ, PK_RADIUS,
  PK_MASS, PK_XX, PK_XY, PK_XZ,
  PK_YY, PK_YZ, PK_ZZ, PK_XXX,
  PK_XYY, PK_XXY, PK_YYY, PK_XXZ,
  PK_YYZ, PK_XYZ, PK_XZZ, PK_YZZ,
  PK_XXXX, PK_XYYY, PK_XXXY, PK_YYYY,
  PK_XXXZ, PK_YYYZ, PK_XXYY, PK_XXYZ,
  PK_XYYZ, PK_XY3S, PK_SOFT, PK_PAD,
  PK_N
};

constexpr int kCellReals = PK_N;                       /* 32 */
constexpr int kCellBytes = kCellReals * sizeof(real);  /* 128 or 256 */
constexpr int kCellPieces = kCellBytes / 16;           /* 8 or 16   */
constexpr int kPieceReals = 16 / sizeof(real);         /* 4 or 2    */
constexpr int kPartReals = 8;
constexpr int kPartBytes = kPartReals * sizeof(real);  /* 32 or 64  */

struct __align__(16) PackedCell { real v[kCellReals]; };
struct __align__(16) PackedPart { real x, y, z, mass, soft, pad0, pad1, pad2; };

static_assert(sizeof(PackedCell) == kCellBytes, "PackedCell size");
static_assert(sizeof(PackedPart) == kPartBytes, "PackedPart size");

/* periodic replica code carried by every list entry:
 * bits 22-24 x+3, 25-27 y+3, 28-30 z+3 */
__host__ __device__ inline int replica_x(int code) { return ((code >> 22) & 7) - 3; }
__host__ __device__ inline int replica_y(int code) { return ((code >> 25) & 7) - 3; }
__host__ __device__ inline int replica_z(int code) { return ((code >> 28) & 7) - 3; }

}  // namespace cb200
#endif
