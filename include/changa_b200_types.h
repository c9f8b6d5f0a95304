/* changa_b200_types.h -- plain-old-data records that cross the host<->GPU seam.
 *
 * Every record here is byte-for-byte the record ChaNGa's Charm++ host code
 * fills in before it calls the GPU gravity entry points, under ChaNGa's
 * default build flags (HEXADECAPOLE, CUDA_2D_TB_KERNEL, no
 * GPU_LOCAL_TREE_WALK).  The struct *tags* are the reference's too, because
 * the entry points have C++ linkage and the tags are part of their mangled
 * names.  Layout authority (reference file:line):
 *
 *   cudatype                 cuda_typedef.h:12      (float; see CUDA_USE_DOUBLE)
 *   CudaVector3D             cuda_typedef.h:55-78
 *   CudaMultipoleMoments     cuda_typedef.h:104-177 (27 reals, hexadecapole)
 *   ILPart / ILCell          cuda_typedef.h:208-229
 *   CompactPartData          cuda_typedef.h:240-243
 *   VariablePartData         cuda_typedef.h:268-272
 *   CudaRequest / CudaDevPtr HostCUDA.h:31-97
 *   EwtData, MomcData, MultipoleMomentsData, EwaldReadOnlyData, EwaldData
 *                            EwaldCUDA.h:11-57
 *
 * CUDA_USE_DOUBLE is OUR addition (the reference hard-wires float): it
 * widens cudatype to double for an FP64 device path; the float layout is the
 * drop-in one.
 */
#ifndef CHANGA_B200_TYPES_H
#define CHANGA_B200_TYPES_H

#include <stddef.h>
#include <cuda_runtime.h>

#ifdef CUDA_USE_DOUBLE
typedef double cudatype;
#else
typedef float cudatype;
#endif

#ifndef HEXADECAPOLE
#define HEXADECAPOLE 1 /* only the hexadecapole layout is built */
#endif

/* ---- geometry ---------------------------------------------------------- */
typedef struct CudaVector3D {
  cudatype x, y, z;
} CudaVector3D;

/* ---- one tree cell: scaled, reduced (trace-free) moments ---------------- */
typedef struct CudaMultipoleMoments {
  cudatype radius;    /* scale length u of the FMOMR components         */
  cudatype soft;      /* mass-weighted softening of the cell            */
  cudatype totalMass; /* monopole                                       */
  CudaVector3D cm;    /* expansion centre                               */
  /* quadrupole (5), octupole (7), hexadecapole (9): order is the ABI.  */
  cudatype xx, xy, xz, yy, yz;
  cudatype xxx, xyy, xxy, yyy, xxz, yyz, xyz;
  cudatype xxxx, xyyy, xxxy, yyyy, xxxz, yyyz, xxyy, xxyz, xyyz;
} CudaMultipoleMoments;

/* ---- interaction-list entries ------------------------------------------ */
/* bucket entry as TreeWalk emits it (expanded to per-particle ILCell by the
 * host serializer, Compute.cpp:1174-1187) */
typedef struct ILPart {
  int index; /* first particle of the source bucket in the device array */
  int off;   /* periodic replica code, bits 22..30                      */
  int num;   /* particles in the source bucket                          */
} ILPart;

/* what the kernels actually stream: one cell, or one source particle */
typedef struct ILCell {
  int index;    /* row in the moments (p-c) or particle (p-p) array     */
  int offsetID; /* bits 22-24 x+3, 25-27 y+3, 28-30 z+3; low 22 ignored */
} ILCell;

/* ---- particles ----------------------------------------------------------- */
typedef struct CompactPartData {
  cudatype mass;
  cudatype soft;
  CudaVector3D position;
} CompactPartData;

typedef struct VariablePartData {
  CudaVector3D a;     /* += acceleration             */
  cudatype potential; /* += potential                */
  cudatype dtGrav;    /* = max( (m_i+m_j)/r^3 ... )  */
} VariablePartData;

/* ---- one offload request (HostCUDA.h:31-89) ------------------------------ */
typedef struct _CudaRequest {
  cudaStream_t stream; /* owned by the caller (DataManager)              */

  CudaMultipoleMoments *d_localMoments;
  CudaMultipoleMoments *d_remoteMoments;
  CompactPartData *d_localParts;
  CompactPartData *d_remoteParts;
  VariablePartData *d_localVars;
  size_t sMoments;
  size_t sCompactParts;
  size_t sVarParts;

  void *list;          /* ILCell[numInteractions], caller-pinned         */
  int *bucketMarkers;  /* [numBucketsPlusOne] offsets into list          */
  int *bucketStarts;   /* [numBucketsPlusOne-1] first target particle    */
  int *bucketSizes;    /* [numBucketsPlusOne-1] target particles         */
  int numInteractions;
  int numBucketsPlusOne;
  void *tp;            /* opaque: requesting TreePiece                   */
  void *missedNodes;   /* remote-resume: moments travelling with request */
  void *missedParts;   /* remote-resume: particles travelling with it    */
  size_t sMissed;      /* bytes in missedNodes / missedParts             */

  int *affectedBuckets; /* opaque to us                                  */
  void *cb;             /* completion callback token                     */
  void *state;          /* opaque                                        */
  cudatype fperiod;     /* one period, applied on all three axes         */

  bool node;   /* bookkeeping only */
  bool remote; /* bookkeeping only */
} CudaRequest;

typedef struct _CudaDevPtr {
  void *d_list;
  int *d_bucketMarkers;
  int *d_bucketStarts;
  int *d_bucketSizes;
} CudaDevPtr;

/* ---- Ewald (EwaldCUDA.h:11-57) ------------------------------------------- */
#define NEWH 80 /* reference's h-table capacity (EwaldCUDA.h:6) */

typedef struct {
  cudatype hx, hy, hz;
  cudatype hCfac, hSfac;
} EwtData;

typedef struct {
  cudatype totalMass;
  cudatype cmx, cmy, cmz;
} MultipoleMomentsData;

typedef struct {
  cudatype m;
  cudatype xx, yy, xy, xz, yz;
  cudatype xxx, xyy, xxy, yyy, xxz, yyz, xyz;
  cudatype xxxx, xyyy, xxxy, yyyy, xxxz, yyyz, xxyy, xxyz, xyyz;
  cudatype zz;
  cudatype xzz, yzz, zzz;
  cudatype xxzz, xyzz, xzzz, yyzz, yzzz, zzzz;
} MomcData;

typedef struct {
  MultipoleMomentsData mm;
  MomcData momcRoot;
  int n, nReps, nEwReps, nEwhLoop;
  cudatype L, fEwCut, alpha, alpha2, k1, ka, fEwCut2, fInner2;
} EwaldReadOnlyData;

typedef struct {
  int EwaldRange[2];             /* small phase only: first/last particle */
  int *EwaldMarkers;             /* large phase: active particle indices  */
  EwtData *ewt;                  /* h-loop table                          */
  EwaldReadOnlyData *cachedData; /* root moments + constants              */
} EwaldData;

/* ---- the layout contract -------------------------------------------------- */
#if defined(__cplusplus)
#define CB200_SASSERT(c, m) static_assert(c, m)
#else
#define CB200_SASSERT(c, m) _Static_assert(c, m)
#endif

CB200_SASSERT(sizeof(CudaVector3D) == 3 * sizeof(cudatype), "CudaVector3D");
CB200_SASSERT(sizeof(CudaMultipoleMoments) == 27 * sizeof(cudatype), "CudaMultipoleMoments");
CB200_SASSERT(sizeof(CompactPartData) == 5 * sizeof(cudatype), "CompactPartData");
CB200_SASSERT(sizeof(VariablePartData) == 5 * sizeof(cudatype), "VariablePartData");
CB200_SASSERT(sizeof(ILCell) == 8, "ILCell");
CB200_SASSERT(sizeof(ILPart) == 12, "ILPart");
CB200_SASSERT(sizeof(EwtData) == 5 * sizeof(cudatype), "EwtData");
CB200_SASSERT(sizeof(MomcData) == 32 * sizeof(cudatype), "MomcData");
CB200_SASSERT(sizeof(MultipoleMomentsData) == 4 * sizeof(cudatype), "MultipoleMomentsData");
#ifndef CUDA_USE_DOUBLE
CB200_SASSERT(sizeof(CudaMultipoleMoments) == 108, "108-byte moments");
CB200_SASSERT(sizeof(CompactPartData) == 20, "20-byte particle core");
CB200_SASSERT(sizeof(VariablePartData) == 20, "20-byte particle vars");
CB200_SASSERT(sizeof(EwaldReadOnlyData) == 192, "192-byte Ewald constants");
CB200_SASSERT(offsetof(CudaRequest, list) == 72, "CudaRequest.list");
CB200_SASSERT(offsetof(CudaRequest, fperiod) == 168, "CudaRequest.fperiod");
CB200_SASSERT(sizeof(CudaRequest) == 176, "CudaRequest");
#endif
CB200_SASSERT(sizeof(CudaDevPtr) == 32, "CudaDevPtr");
CB200_SASSERT(sizeof(EwaldData) == 32, "EwaldData");

#endif /* CHANGA_B200_TYPES_H */
