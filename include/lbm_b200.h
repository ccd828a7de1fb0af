/*
 * lbm_b200.h -- C ABI of liblbm_b200.so: the D3Q19 alpha/beta lattice-Boltzmann time step,
 * its boundary/field access and its halo pack/unpack as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary that replaces the reference's OpenCL plumbing
 * (src/libcl/CCL.hpp: CContext/CCommandQueue/CMem/CProgram/CKernel) underneath
 * CLbmSolver<T>.  Every entry point names the reference interface it replaces
 * (paths relative to the reference tree).  Plain C types only: no C++ exceptions, no torch
 * types and no CUDA types cross this boundary (streams and device buffers travel as void*).
 *
 * Conventions
 *   - every function returns an int status, LBM_OK (0) on success; the message of the last
 *     failure of a handle is available through lbmGetLastErrorString (handle == NULL: the
 *     last failure of lbmCreate / handle-less calls on this thread);
 *   - host pointers are caller-owned and are fully consumed/produced before return
 *     (the reference's blocking enqueueReadBuffer, src/CLbmSolver.hpp:715-716);
 *   - the library owns all device memory; every call selects the handle's device itself,
 *     so one host thread may drive several devices and several host threads may each
 *     drive their own handle;
 *   - origin/size triples are (x, y, z) in cells of the sub-domain *including* its ghost
 *     layers, packed buffers are [component][z][y][x] (src/CLbmSolver.hpp:706-710).
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_OK                 0
#define LBM_ERR_INVALID        1   /* bad argument */
#define LBM_ERR_CUDA           2   /* CUDA runtime failure (message has the CUDA error) */
#define LBM_ERR_NO_DEVICE      3   /* no CUDA device: the product has no CPU fallback */
#define LBM_ERR_TIMEOUT        4   /* a device-side halo wait gave up: the neighbour never pushed its face
                                      (reported by lbmWait; LBM_B200_WAIT_TIMEOUT_MS, default 30000) */

/* cell flags, src/common.h:19-22 */
#define LBM_FLAG_OBSTACLE            (1 << 0)
#define LBM_FLAG_FLUID               (1 << 1)
#define LBM_FLAG_VELOCITY_INJECTION  (1 << 2)
#define LBM_FLAG_GHOST_LAYER         (1 << 3)

#define LBM_F32 0
#define LBM_F64 1

#define LBM_SIZE_DD_HOST 19          /* CLbmSolver::SIZE_DD_HOST, src/CLbmSolver.hpp:72 */

/* beta accumulation order (src/cl_programs/lbm_beta.cl) */
#define LBM_BETA_ORDER_SHIPPED 0     /* shared-memory path, the kernel as shipped (:256-483) */
#define LBM_BETA_ORDER_LINEAR  1     /* the reference's USE_SHARED_MEMORY 0 path (:53-164)    */

/* halo payload */
#define LBM_HALO_SLOTS_REFERENCE 0   /* all 19 slots, what the reference ships */
#define LBM_HALO_SLOTS_MINIMAL   1   /* the 5 slots a neighbour consumes        */

#define LBM_SYNC_ALPHA 0             /* CController::syncAlpha, src/CController.hpp:265-320 */
#define LBM_SYNC_BETA  1             /* CController::syncBeta,  src/CController.hpp:322-383 */

typedef struct lbm_solver *lbm_t;

/*
 * Everything CLbmSolver's constructor + reload() bake into the kernels
 * (src/CLbmSolver.hpp:221-260, 319-373, 583-614).  Floating-point members carry values of
 * the simulation type T widened to double (exact for float).
 */
typedef struct lbm_desc {
	uint32_t struct_size;        /* = sizeof(lbm_desc), ABI guard */
	int32_t  device;             /* CUDA device ordinal (conf.xml device-number) */
	int32_t  dtype;              /* LBM_F32 | LBM_F64: the reference's `typedef float T` */
	int32_t  size[3];            /* sub-domain cells incl. ghost layers (DOMAIN_CELLS_X/Y/Z) */
	int32_t  bc[6];              /* face flags x0,x1,y0,y1,z0,z1 (init_kernel bc[], :396-403) */
	double   inv_tau;            /* kernel arg 4 */
	double   tau;                /* only read when smagorinsky_cs != 0 */
	double   gravitation[3];     /* kernel args 5-7 (lattice units) */
	double   u_lid;              /* kernel arg 8: drivenCavityVelocity[0]*d_timestep */
	int32_t  store_velocity;     /* STORE_VELOCITY */
	int32_t  store_density;      /* STORE_DENSITY */
	double   smagorinsky_cs;     /* 0 = plain BGK (the reference); >0 = LES extension */
	int32_t  beta_order;         /* LBM_BETA_ORDER_* */
	int32_t  work_group_size;    /* reference `kernel-count` (128): only selects the
	                                work-group x-shift of lbm_beta.cl:221-234; 0 = off */
	int32_t  block_size;         /* CUDA block size hint, 0 = library default */
	int32_t  vector_width;       /* cells per thread, 0 = widest that divides size[0] */
	void    *compute_stream;     /* optional external cudaStream_t for the step kernels */
	void    *comm_stream;        /* optional external cudaStream_t for halo kernels */
} lbm_desc;

/* ---- bring-up: replaces CCL::CPlatforms/CContext/CDevices/CCommandQueue
 *      (src/CController.hpp:83-226) ------------------------------------------------- */
int lbmGetDeviceCount(int *count);
int lbmGetVersion(void);
const char *lbmGetLastErrorString(lbm_t h);

/* ---- CLbmSolver::CLbmSolver + reload() (src/CLbmSolver.hpp:221-260,272-617) and the
 *      destructor of its CMem/CKernel members ------------------------------------------ */
int lbmCreate(lbm_t *out, const lbm_desc *desc);
int lbmDestroy(lbm_t h);

/* ---- CLbmSolver::reset / simulationStep / simulationStepAlpha / simulationStepBeta /
 *      wait / simulation_step_counter (src/CLbmSolver.hpp:619-683) ---------------------- */
int lbmReset(lbm_t h);
int lbmStep(lbm_t h);                 /* counter&1 ? alpha : beta; counter++ */
int lbmStepAlpha(lbm_t h);
int lbmStepBeta(lbm_t h);
int lbmSteps(lbm_t h, int nsteps);    /* nsteps x lbmStep without host round trips */
int lbmWait(lbm_t h);                 /* CCommandQueue::finish */
int lbmGetStepCounter(lbm_t h, uint64_t *counter);
int lbmSetStepCounter(lbm_t h, uint64_t counter);

/* CLbmSolver::addDrivenCavityValue (src/CLbmSolver.hpp:262-270): new kernel arg 8 */
int lbmSetDrivenCavityVelocity(lbm_t h, double u_lid);

/* ---- field access: CLbmSolver::store* / set* (src/CLbmSolver.hpp:688-978).
 *      origin == NULL: the whole array (the non-rect overloads). ------------------------ */
int lbmStoreDD(lbm_t h, void *host_dst, const int origin[3], const int size[3]);
/* norm == NULL: all 19 slots (:719-735); else only slots with norm . e_f > 0 (:737-757) */
int lbmSetDD(lbm_t h, const void *host_src, const int origin[3], const int size[3], const int norm[3]);
int lbmStoreVelocity(lbm_t h, void *host_dst, const int origin[3], const int size[3]);
int lbmSetVelocity(lbm_t h, const void *host_src, const int origin[3], const int size[3]);
int lbmStoreDensity(lbm_t h, void *host_dst, const int origin[3], const int size[3]);
int lbmSetDensity(lbm_t h, const void *host_src, const int origin[3], const int size[3]);
int lbmStoreFlags(lbm_t h, int *host_dst, const int origin[3], const int size[3]);
int lbmSetFlags(lbm_t h, const int *host_src, const int origin[3], const int size[3]);

/* CLbmSolver::getVelocityChecksum (src/CLbmSolver.hpp:1103-1123).
 * host_order != 0: the reference's serial float accumulation in index order (exact);
 * host_order == 0: device reduction (warp shuffles + one atomic per block) in double. */
int lbmChecksumVelocity(lbm_t h, double *out, int host_order);

/* ---- device-resident halo path: replaces the host-staged bodies of
 *      CController::syncAlpha/syncBeta (src/CController.hpp:265-383), i.e.
 *      storeDensityDistribution -> MPI -> setDensityDistribution.  The rect/direction
 *      arguments are exactly the CComm fields (src/CComm.hpp:8-79, values from
 *      src/CManager.hpp:122-199).  dev_buf is DEVICE memory of lbmHaloBytes bytes. ------ */
/* The slots a sync moves for a face whose CComm direction ON THE RECEIVING sub-domain is
 * recv_dir (the unit normal pointing into the receiver, src/CManager.hpp:122-199):
 *   LBM_SYNC_BETA : slots with e_f . recv_dir > 0 (what setDensityDistribution(..., norm)
 *                   writes, src/CLbmSolver.hpp:747-754) -- both payload modes;
 *   LBM_SYNC_ALPHA: REFERENCE = all 19 slots (src/CController.hpp:290-313);
 *                   MINIMAL   = slots with e_f . recv_dir < 0, the only ones the next beta
 *                   step of the receiver pulls out of its ghost layer. */
int lbmHaloSlotMask(int sync_kind, const int recv_dir[3], int slots, uint32_t *mask);
int lbmHaloBytes(lbm_t h, const int size[3], uint32_t slot_mask, size_t *bytes);
/* dd rect -> dev_buf, layout [selected slot, ascending][z][y][x]; stream NULL = comm stream */
int lbmHaloPack(lbm_t h, const int origin[3], const int size[3], uint32_t slot_mask,
		void *dev_buf, void *stream);
/* dev_buf (holding buf_slot_mask) -> dd rect, writing only write_mask (subset) */
int lbmHaloUnpack(lbm_t h, const int origin[3], const int size[3], uint32_t buf_slot_mask,
		uint32_t write_mask, const void *dev_buf, void *stream);
/* same-process peers: ONE kernel copies the src rect straight into the dst sub-domain over
 * NVLink peer access (pack + send + unpack fused); runs on `stream` of the SOURCE device
 * (NULL = its comm stream). */
int lbmHaloCopyPeer(lbm_t src, const int src_origin[3], lbm_t dst, const int dst_origin[3],
		const int size[3], uint32_t slot_mask, void *stream);

/* ---- one-sided halo exchange over NVLink peer memory (no NCCL, no host staging) --------
 * A face = one CComm (src/CComm.hpp:8-79).  Each face owns a RECEIVE block in device memory
 * ([flag words][alpha staging][beta staging]); the neighbour maps it (same process: directly;
 * other process: CUDA IPC) and its push kernel packs the face and stores it straight into
 * that block over NVLink, then raises the block's flag word; the owner's pull waits on the
 * flag (device side) and unpacks.  Replaces syncAlpha/syncBeta (src/CController.hpp:265-383).
 * Meant for one rank per GPU.  Several ranks MAY share a GPU (tests do), but a pull kernel spins
 * until the neighbour's push has run, so all streams of the ranks on one device (two per rank) must
 * be able to make progress independently: keep them within the device's hardware queues
 * (CUDA_DEVICE_MAX_CONNECTIONS, default 8 -> at most 4 free-running ranks per GPU), or enqueue every
 * push of an axis before any pull of that axis from one host thread (lbmCommPush / lbmCommPull). */
int lbmCommAddFace(lbm_t h, int dst_rank, const int send_origin[3], const int recv_origin[3],
		const int size[3], const int dir[3], int slots, int *face_id);
int lbmCommFaceCount(lbm_t h, int *count);
int lbmCommGetIpcHandle(lbm_t h, int face_id, void *handle64);            /* 64-byte cudaIpcMemHandle_t */
int lbmCommConnectIpc(lbm_t h, int face_id, const void *peer_handle64);   /* peer in another process */
int lbmCommConnectLocal(lbm_t h, int face_id, lbm_t peer, int peer_face_id); /* peer in this process */
/* Sequence numbers: every push of a face counts up a device-resident word and publishes it in the
 * neighbour's flag; every pull counts up its own word and waits for the flag to reach it.  Nothing about
 * them lives on the host or in kernel arguments, so a CUDA graph captured around lbmCommStep (or around
 * push/pull calls) can be replayed any number of times and keeps synchronising.  Each face must see
 * exactly one push and one pull per sync kind and step, in step order -- lbmCommSync / lbmCommStep do. */
int lbmCommBeginSync(lbm_t h, int sync_kind);           /* kept for source compatibility: validates sync_kind, nothing else */
int lbmCommPush(lbm_t h, int sync_kind, int axis);      /* push my faces of one axis (comm stream) */
int lbmCommPull(lbm_t h, int sync_kind, int axis);      /* wait (one device thread per face) + unpack my faces of one axis */
int lbmCommSync(lbm_t h, int sync_kind);                /* for each axis in order: push; pull */
/* Order of the three axis phases of a sync.  Any fixed order delivers the same halo (every face
 * spans the full extent of the other two axes, so a later phase forwards the rims an earlier one
 * received); only the leftovers in ghost cells nobody reads differ.
 *   LBM_AXIS_ORDER_XYZ  the reference's CComm walk (src/CManager.hpp:122-199), the default;
 *   LBM_AXIS_ORDER_ZYX  z, y, then x.  lbmCommStep then keeps the z and y faces under the
 *                       interior kernel and does not split an x shell off (it would touch both ends
 *                       of every row of the sub-domain -- one DRAM page per row and slot -- and cost
 *                       half a step whatever its width).  Instead the x faces leave the step kernels
 *                       themselves: the threads that own the cells next to an x ghost face store the
 *                       5 populations the neighbour consumes straight into its receive block (peer
 *                       stores from registers, no strided gather afterwards); behind the step kernel
 *                       a small rim pass forwards the edge lines the y/z phases delivered and raises
 *                       the flag, then wait + unpack.  Use it when the decomposition cuts x.
 *                       (5-slot payload only, rows that are a whole number of thread blocks of 128 or
 *                       of a smaller multiple of 32; otherwise, or with LBM_B200_XFUSE=0, the x faces go
 *                       through separate push / wait / unpack kernels behind the step kernel.)
 * Every rank of a run must use the same order.  LBM_B200_AXIS_ORDER=xyz|zyx presets it. */
#define LBM_AXIS_ORDER_XYZ 0
#define LBM_AXIS_ORDER_ZYX 1
int lbmCommSetAxisOrder(lbm_t h, int order);
int lbmCommGetAxisOrder(lbm_t h, int *order);
/* one overlapped time step: shell kernels -> (push/pull on the comm stream || interior kernel)
 * -> join; == CController::computeNextStep (src/CController.hpp:385-391) */
int lbmCommStep(lbm_t h);
/* the same step with CUDA events on both streams; ms = time from the fork to {shell kernels done,
 * interior kernel done, exchange (push/wait/unpack) done, join}: shows how much of the halo
 * exchange is hidden under the interior kernel.  Synchronises. */
int lbmCommStepTimed(lbm_t h, float ms[4]);

/* ---- overlap support: the step split into the shell next to ghost faces and the interior.
 *      ghost_faces: bit a*2+s set = face (axis a, side s) has a neighbour. --------------- */
int lbmStepShell(lbm_t h, int ghost_faces);     /* launches on the compute stream */
int lbmStepShellComm(lbm_t h, int ghost_faces); /* same, on the comm stream (runs next to the interior) */
int lbmStepInterior(lbm_t h, int ghost_faces);  /* launches on the compute stream; counter++ */
int lbmStreamWaitStream(lbm_t h, int waiter_is_comm); /* event edge between the two streams */
int lbmGetStreams(lbm_t h, void **compute_stream, void **comm_stream);

/* ---- raw device pointers for zero-copy interop (dd, flags, velocity, density).
 *      dd is [19][slot_stride] with slot_stride >= cells (lbmGetSlotStride; equal to cells
 *      unless the tuning hook LBM_B200_SLOT_PAD_BYTES pads it); the host-buffer entry points
 *      above always use the dense [19][cells] layout of the reference. ---------------------- */
#define LBM_BUF_DD 0
#define LBM_BUF_FLAGS 1
#define LBM_BUF_VELOCITY 2
#define LBM_BUF_DENSITY 3
int lbmGetDevicePointer(lbm_t h, int which, void **ptr, size_t *bytes);
int lbmGetSlotStride(lbm_t h, size_t *cells);

/* ---- timing on the compute stream with CUDA events (replaces CStopwatch around the loop,
 *      src/CController.hpp:429-438) --------------------------------------------------- */
int lbmTimerStart(lbm_t h);
int lbmTimerStop(lbm_t h, float *milliseconds);   /* synchronises */

/* number of kernel launches issued by this handle since creation (bench gpu_launches) */
int lbmGetLaunchCount(lbm_t h, uint64_t *launches);

/* ---- per-kernel device timeline (replaces the reference's PROFILE build: the OpenCL queue's
 *      CL_QUEUE_PROFILING_ENABLE + one CProfilerEvent per enqueue, src/libcl/CCL.hpp:1244-1245,
 *      1752-1778; src/libtools/CProfilerEvent.hpp:29-38, CProfiler.hpp:35-48).
 *      LBM_PROFILE_EVENTS: every kernel launch of this handle is bracketed by CUDA events on its
 *      stream (nothing blocks; the reference waits after each enqueue); LBM_PROFILE_NVTX: every
 *      launch sits in an NVTX range.  Names are the reference's kernel names -- init_kernel,
 *      lbm_kernel_alpha, lbm_kernel_beta, copy_buffer_rect -- plus lbm_kernel_beta.wrap,
 *      halo_push, halo_pull, checksum_kernel for the kernels the reference does not have.
 *      Times are nanoseconds since lbmProfileEnable / lbmProfileClear.  The environment
 *      variable LBM_B200_PROFILE=<mode> enables it from lbmCreate on (init_kernel included).
 *      Not usable while the launches are being captured into a CUDA graph. ---------------- */
#define LBM_PROFILE_EVENTS 1
#define LBM_PROFILE_NVTX 2
int lbmProfileEnable(lbm_t h, int mode);          /* 0 = off */
int lbmProfileClear(lbm_t h);                     /* synchronises, drops the recorded events, new time zero */
int lbmProfileEventCount(lbm_t h, uint64_t *count, uint64_t *dropped /* may be NULL */);
int lbmProfileGetEvent(lbm_t h, uint64_t index, char *name, size_t name_bytes,
		uint64_t *start_ns, uint64_t *end_ns);    /* synchronises on that event */

/* resolved launch configuration (cells per thread, block size, active work-group quirk) */
int lbmGetConfig(lbm_t h, int *vector_width, int *block_size, int *wg_quirk);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
