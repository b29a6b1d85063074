/*
 * dynsight_b200.h -- C ABI of the B200-native descriptor engine (libdynsight_b200.so).
 *
 * Drop-in boundary for ONE hot path of GMPavanLab/dynsight: per-frame periodic neighbour search
 * -> LENS / coordination number, SOAP power spectrum -> timeSOAP.  The reference is pure Python
 * (numba + DScribe); this header is what its FFI for the path would bind (ctypes stubs are shown
 * in INTEGRATION.md).  Each entry point cites the reference interface it replaces
 * (paths relative to /root/reference/src/dynsight/_internal/).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer (owned by the caller, e.g. a torch tensor);
 *     h_* is a HOST pointer.  No allocation happens inside the library except tiny pinned staging.
 *   - all arrays are dense, C-contiguous; frames are the slowest axis of every batched array.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously on it.
 *   - return value: 0 = ok, negative = error (DSG_ERR_*); dsg_last_error() gives the message of
 *     the calling thread's last failure.  Nothing is thrown across the boundary.
 *   - thread safety: no global mutable state; calls on distinct streams may run concurrently.
 *   - pos_dtype: DSG_F32 (what MDAnalysis hands out) or DSG_F64 (what the numba kernels take).
 */
#ifndef DYNSIGHT_B200_H
#define DYNSIGHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSG_OK 0
#define DSG_ERR_INVALID_ARGUMENT (-1)
#define DSG_ERR_CUDA (-2)
#define DSG_ERR_WORKSPACE_TOO_SMALL (-3)
#define DSG_ERR_UNSUPPORTED (-4)
#define DSG_ERR_OVERFLOW (-5)

#define DSG_F32 0
#define DSG_F64 1

/* Library / build identification. */
const char* dsg_version(void);
/* Message of the last error raised on the calling thread ("" if none). */
const char* dsg_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200); <0 on error. */
int dsg_device_arch(void);
/* Number of kernel launches this library has issued in the process (monotonic counter;
 * bench.py differences it around the timed region to report "gpu_launches"). */
long long dsg_launch_count(void);
/* Measurement aid (bench.py): launches `blocks` x 256 threads doing `iters` rounds of 16 independent
 * float64 FMAs each; *flops_out = floating-point operations issued.  d_scratch needs blocks*256
 * doubles.  Timed by the caller with CUDA events to obtain the fp64 FMA peak of the device. */
int dsg_fp64_fma_probe(double* d_scratch, int blocks, int iters, void* stream, double* flops_out);

/* ------------------------------------------------------------------------------------------
 * Neighbour lists (replaces lens/lens.py:29-147: build_cell_list +
 * neighbor_list_celllist_centers, batched over frames).
 *
 * Membership rule reproduced bit-exactly (lens.py:98-104): env atom j is a neighbour of centre i
 * iff  j != i  (env-local vs centre-local index)  and  |dr|^2 < (r_cut - 1e-6)^2  with
 * dr = pos_env[j] - pos_cent[i] evaluated in float64, minimum-imaged per axis iff respect_pbc and
 * box[k] > 0; candidates come from the 27 cells around the centre's cell of the reference's own
 * grid: ncell_k = max(3, int(box_k // r_cut)), cell = int(x / box * ncell) % ncell.
 * Rows are sorted ascending (lens.py:142-145).
 *
 * Two-phase protocol (the caller owns every buffer):
 *   1. dsg_neigh_workspace_bytes(...)                      -> size of the scratch buffer
 *   2. dsg_neigh_count(..., d_workspace, d_counts, d_indptr, d_frame_off)
 *      builds the cell lists of all frames in the workspace, writes per-centre counts,
 *      per-frame-local CSR row pointers and the exclusive int64 frame offsets
 *   3. caller reads d_frame_off[n_frames] (= total nnz), allocates d_indices
 *   4. dsg_neigh_fill(..., d_workspace, d_indptr, d_frame_off, d_indices)
 *      (same arguments as step 2; the workspace must be untouched in between)
 * Row i of frame f is d_indices[d_frame_off[f] + d_indptr[f*(n_cent+1)+i] ...+i+1]).
 *
 * d_pos_cent == NULL means "centres are the env atoms" (centers == selection).
 * h_box: (n_frames, 3) float64 box lengths on the HOST (ts.dimensions[:3], lens.py:268-274).
 * ------------------------------------------------------------------------------------------ */
int dsg_neigh_workspace_bytes(int n_frames, int n_env, int n_cent, int same_centres,
                              const double* h_box, double r_cut, size_t* bytes_out);

int dsg_neigh_count(const void* d_pos_env, const void* d_pos_cent, int pos_dtype, int n_frames,
                    int n_env, int n_cent, const double* h_box, double r_cut, int respect_pbc,
                    void* d_workspace, size_t workspace_bytes, int32_t* d_counts,
                    int32_t* d_indptr, int64_t* d_frame_off, void* stream);

int dsg_neigh_fill(const void* d_pos_env, const void* d_pos_cent, int pos_dtype, int n_frames,
                   int n_env, int n_cent, const double* h_box, double r_cut, int respect_pbc,
                   void* d_workspace, size_t workspace_bytes, const int32_t* d_indptr,
                   const int64_t* d_frame_off, int32_t* d_indices, void* stream);

/* Rows path (LENS fast path): ONE traversal.  compute_lens (lens.py:192-310) never returns the
 * lists, so for LENS the rows need neither canonical CSR order nor sorting.  Membership is decided
 * exactly as above (same sets), but row i of frame f is stored UNSORTED at
 *   d_rows[d_row_start[f*n_cent+i] ... + d_counts[f*n_cent+i])
 * where the start was reserved with an atomic add on one of 256 cursors.  d_status must hold
 * 2+256 int64: after the call status[0] != 0 means that at least one reservation did not fit
 * (those rows have row_start = -1): retry with rows_capacity >= status[1]*max(status[2..258)) * 1.05
 * (status[1] = number of regions actually used, a power of two <= 256).
 * respect_pbc: bit 0 = apply the minimum image; bit 1 = force the warp-per-cell kernel; bit 2 =
 * inclusive cutoff "dr2 <= r_cut^2" (the FastNS rule of analysis/spatial_average.py:39-40)
 * instead of the numba kernels' "dr2 < (r_cut - 1e-6)^2".  Without bit 1 sparse grids use one
 * thread per centre (rows longer than its staging tile are written by a second walk of the same
 * thread).  status[0] == 2 is reserved: a caller seeing it must repeat the call with bit 1 set. */
int dsg_neigh_rows(const void* d_pos_env, const void* d_pos_cent, int pos_dtype, int n_frames,
                   int n_env, int n_cent, const double* h_box, double r_cut, int respect_pbc,
                   void* d_workspace, size_t workspace_bytes, int32_t* d_counts,
                   int64_t* d_row_start, int32_t* d_rows, int64_t rows_capacity,
                   int64_t* d_status, void* stream);

/* Canonical CSR (what dsg_neigh_count + dsg_neigh_fill produce: numba-identical indptr, rows
 * ascending -- lens.py:108-145) from the rows of dsg_neigh_rows, i.e. with ONE traversal of the
 * cell list instead of two:
 *   dsg_rows_csr_offsets: indptr (n_frames, n_cent+1) frame-local exclusive scan of the counts and
 *     frame_off (n_frames+1); the caller reads frame_off[n_frames] (nnz) and allocates d_indices;
 *   dsg_rows_to_csr: every row is copied to indices[frame_off[f] + indptr[f][i] ...] and sorted;
 *     expected_row_len (0 = unknown) picks one thread per row (<= 24) or one warp per row. */
int dsg_rows_csr_workspace_bytes(int n_frames, int n_cent, size_t* bytes_out);

int dsg_rows_csr_offsets(const int32_t* d_counts, int n_frames, int n_cent, void* d_workspace,
                         size_t workspace_bytes, int32_t* d_indptr, int64_t* d_frame_off,
                         void* stream);

int dsg_rows_to_csr(const int32_t* d_counts, const int64_t* d_row_start, const int32_t* d_rows,
                    const int32_t* d_indptr, const int64_t* d_frame_off, int n_frames, int n_cent,
                    int32_t* d_indices, int expected_row_len, void* stream);

/* LENS over the rows of dsg_neigh_rows (same output contract as dsg_lens_pairs).  For delay 1 one
 * thread owns one centre and a run of pairs (both rows of a pair in registers; warps that meet a
 * row longer than 32 switch to the warp-cooperative intersection).  Bit 30 of `delay` = rows are
 * expected to be long (more than ~24 entries): one warp per centre from the start. */
int dsg_lens_pairs_rows(const int32_t* d_counts, const int64_t* d_row_start,
                        const int32_t* d_rows, int n_frames, int n_cent, int delay, double* d_out,
                        int64_t ld_out, int64_t col0, void* stream);

/* Pair path: coordination numbers and LENS of a batch of frames WITHOUT neighbour lists.
 * Replaces, for centers == selection (the default of compute_lens, lens.py:192-199), the whole
 * chain neighbor_list_celllist_centers (lens.py:69-147) x 2 + lens_from_two_csr (lens.py:150-189)
 * of one frame pair, and the counting loop of Trj.get_coord_number (trajectory.py:168-184).
 * The neighbour relation of the reference is symmetric when centres and environment coincide,
 * so every pair is tested once (half of the 27 cells) and credited to both atoms; a pair found
 * in frame t is looked up directly in frame t + delay.  Same membership rule, bit for bit:
 * dr2 < (r_cut - 1e-6)^2 on the minimum image (if respect_pbc), cells of the reference grid,
 * atoms outside [0, box) binned by truncation + modulo like lens.py:52-54.
 *   d_counts (n_frames, n_atoms) int32 or NULL;
 *   d_lens or NULL: d_lens[i*lens_stride + col0 + t], t in [0, n_frames - delay), float64.
 * dsg_lens_direct_supported: 1 if this path applies to the batch (sparse cells, i.e. one thread
 * per atom is the right granularity; at least six reference cells per periodic axis), 0 if the
 * caller has to use dsg_neigh_rows + dsg_lens_pairs_rows, < 0 on invalid arguments. */
int dsg_lens_direct_supported(int n_frames, int n_atoms, const double* h_box, double r_cut,
                              int respect_pbc);

int dsg_lens_direct_workspace_bytes(int n_frames, int n_atoms, const double* h_box, double r_cut,
                                    size_t* bytes_out);

int dsg_lens_direct(const void* d_pos, int pos_dtype, int n_frames, int n_atoms,
                    const double* h_box, double r_cut, int respect_pbc, int delay,
                    void* d_workspace, size_t workspace_bytes, int32_t* d_counts, double* d_lens,
                    int64_t lens_stride, int64_t col0, void* stream);

/* ------------------------------------------------------------------------------------------
 * LENS (replaces lens/lens.py:150-189 lens_from_two_csr, batched over frame pairs).
 *
 * Frames a = 0..n_pairs-1 are paired with b = a + delay inside one CSR batch as produced above
 * (n_frames = n_pairs + delay).  out[i*ld_out + col0 + a] = ((|A|+|B|) - 2|A n B|) / (|A|+|B|)
 * in IEEE float64, 0.0 when |A|+|B| == 0 -- the (n_centers, n_pairs) layout of compute_lens
 * (lens.py:262,303).
 * ------------------------------------------------------------------------------------------ */
int dsg_lens_pairs(const int32_t* d_indptr, const int64_t* d_frame_off, const int32_t* d_indices,
                   int n_frames, int n_cent, int delay, double* d_out, int64_t ld_out,
                   int64_t col0, void* stream);

/* Single pair of independent CSR lists: exactly the signature of lens_from_two_csr. */
int dsg_lens_from_two_csr(const int32_t* d_indptr1, const int32_t* d_indices1,
                          const int32_t* d_indptr2, const int32_t* d_indices2, int n_cent,
                          double* d_out, void* stream);

/* Coordination numbers (replaces trajectory/trajectory.py:168-184): counts (n_frames, n_cent)
 * int32 -> out[i*ld_out + col0 + f] float64, the (n_centers, n_frames) layout of the Insight. */
int dsg_counts_to_f64(const int32_t* d_counts, int n_frames, int n_cent, double* d_out,
                      int64_t ld_out, int64_t col0, void* stream);

/* ------------------------------------------------------------------------------------------
 * Neighbour-list consumers (replace descriptors/misc.py:15-219).
 * Both take the canonical CSR of dsg_neigh_count/fill with centres == selection == all atoms.
 *
 * dsg_orientational_op: |psi_n| of misc.py:76-93 -- out[i*ld_out + col0 + f] =
 *   |sum_{j in row, j != i} e^{i n theta_ij}| / len(row), 0 if len(row) <= 1; theta from the raw
 *   (x, y) differences (no minimum image, z ignored), like the reference.
 * dsg_velocity_alignment: phi of misc.py:98-219 -- mean cosine similarity between the vector of
 *   atom i and those of its neighbours (zero vectors and j == i skipped).  displacement != 0:
 *   d_vec holds positions (n_out+1, n_atoms, 3) and the vector of output f is
 *   pos[f+1] - pos[0] (the reference never advances r_0, misc.py:207-213) with the rows of frame
 *   f; displacement == 0: d_vec holds (n_out, n_atoms, 3) velocities, rows of the same frame.
 * ------------------------------------------------------------------------------------------  * expected_row_len (0 = unknown): mean neighbours per atom; 4, 8, 16 or 32 lanes share one
 * (atom, frame) accordingly.
 */
int dsg_orientational_op(const void* d_pos, int pos_dtype, const int32_t* d_indptr,
                         const int64_t* d_frame_off, const int32_t* d_indices, int n_frames,
                         int n_atoms, int order, double* d_out, int64_t ld_out, int64_t col0,
                         int expected_row_len, void* stream);

int dsg_velocity_alignment(const void* d_vec, int dtype, int displacement,
                           const int32_t* d_indptr, const int64_t* d_frame_off,
                           const int32_t* d_indices, int n_out, int n_atoms, double* d_out,
                           int64_t ld_out, int64_t col0, int expected_row_len, void* stream);

/* ------------------------------------------------------------------------------------------
 * timeSOAP (replaces timesoap/timesoap.py:13-179: normalize_soap + soap_distance + timesoap).
 *
 * soap is addressed as soap[i*stride_i + f*stride_f + c] (element strides, float64), which covers
 * both the contiguous (n_particles, n_frames, n_comp) array and the transposed view
 * saponify_trajectory returns (saponify.py:124).  out[i*ld_out + col0 + f] for f in
 * [0, n_frames-delay):  sqrt(max(0, 2 - 2*clip(v1.v2/(|v1||v2|), -1, 1))), with the reference's
 * convention that a zero-norm vector normalises to the zero vector (result sqrt(2)).
 * Returns DSG_ERR_INVALID_ARGUMENT if delay < 1 or delay >= n_frames (timesoap.py:175-177).
 * ------------------------------------------------------------------------------------------ */
int dsg_timesoap(const double* d_soap, int n_particles, int n_frames, int n_comp,
                 int64_t stride_i, int64_t stride_f, int delay, double* d_out, int64_t ld_out,
                 int64_t col0, void* stream);

/* normalize_soap (timesoap/timesoap.py:13-58): n_vectors contiguous vectors of n_comp float64,
 * out = v/|v| where |v| > 0 else 0. */
int dsg_normalize_soap(const double* d_in, double* d_out, int64_t n_vectors, int n_comp,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * SOAP power spectrum (replaces soapify/saponify.py:98-122, i.e. DScribe
 * SOAP(species, r_cut, n_max, l_max, periodic, sigma=1, rbf="gto").create, batched over frames).
 *
 * h_alphas (l_max+1, n_max), h_betas (l_max+1, n_max, n_max): GTO exponents and Loewdin
 * orthonormalisation matrices of DScribe's get_basis_gto, computed by the host in float64.
 * d_species (n_atoms) int32 in [0, n_species): index into the ascending-atomic-number species list.
 * d_centres (n_cent) int32: selection-local indices of the centre atoms.
 * h_box (n_frames,3) float64 or NULL (non-periodic / pbc off): orthorhombic cell lengths.
 * Every periodic image within r_cut + sqrt(-2 ln 1e-3) contributes (centre included).
 * out[i*stride_i + f*stride_f + c], c in [0, n_feat) with
 * n_feat = (S*n_max)*(S*n_max+1)/2*(l_max+1); order: species pairs Z<=Z', then l, n, n'
 * (n'>=n when Z==Z').  float64 output, float64 arithmetic.
 * Workspace (dsg_soap_workspace_bytes, same shapes and h_box as the call): the cell-sorted copies,
 * the primitive sums between the two kernels when n_species > 1 or n_max > 8, and -- for periodic
 * frames with at least three search-radius cells per axis -- per-centre neighbour rows of
 * 4 * (1.5 * mean neighbours + 64) bytes per atom, the mean taken from n_atoms / box volume (left
 * out when fewer than a quarter of the atoms are centres and the rows would exceed 256 MiB).  Rows that
 * overflow (strongly clustered input) are handled on the device by the in-kernel search; no status
 * has to be read back.
 * ------------------------------------------------------------------------------------------ */
int dsg_soap_n_features(int n_species, int n_max, int l_max);

int dsg_soap_workspace_bytes(int n_frames, int n_atoms, int n_cent, int n_species, int n_max,
                             int l_max, const double* h_box, double r_cut, size_t* bytes_out);

int dsg_soap(const void* d_pos, int pos_dtype, const int32_t* d_species, const int32_t* d_centres,
             int n_frames, int n_atoms, int n_cent, int n_species, const double* h_box,
             double r_cut, int n_max, int l_max, const double* h_alphas, const double* h_betas,
             void* d_workspace, size_t workspace_bytes, double* d_out, int64_t stride_i,
             int64_t stride_f, void* stream);

/* SOAP fused with timeSOAP (replaces the chain Trj.get_soap -> Insight.get_angular_velocity ->
 * timesoap, trajectory.py:267-320 / insight.py:136-155 / timesoap.py:130-179, for callers that want
 * timeSOAP only): the spectrum of frame f is folded into the distance to the spectrum of frame
 * f - delay inside the spectrum epilogue.  A warp owns one centre and a run of frame pairs; the
 * `delay` previous vectors live in its ring (a workspace region rewritten every `delay` frames, i.e.
 * L2 traffic), so the (n_cent, n_frames, n_feat) spectra never exist:
 *   d_timesoap[i*ts_ld + ts_col0 + t] = timesoap value of the pair (t, t + delay), t in
 *   [0, n_frames - delay), float64 -- the (n_centers, n_frames - delay) layout of timesoap().
 * d_soap_out may be NULL; if given, the spectra are written as well (same layout as dsg_soap).
 * Same arithmetic as dsg_soap followed by dsg_timesoap (equal to rounding). */
int dsg_soap_timesoap_workspace_bytes(int n_frames, int n_atoms, int n_cent, int n_species,
                                      int n_max, int l_max, const double* h_box, double r_cut,
                                      int delay, size_t* bytes_out);

int dsg_soap_timesoap(const void* d_pos, int pos_dtype, const int32_t* d_species,
                      const int32_t* d_centres, int n_frames, int n_atoms, int n_cent,
                      int n_species, const double* h_box, double r_cut, int n_max, int l_max,
                      const double* h_alphas, const double* h_betas, int delay, void* d_workspace,
                      size_t workspace_bytes, double* d_timesoap, int64_t ts_ld, int64_t ts_col0,
                      double* d_soap_out, int64_t stride_i, int64_t stride_f, void* stream);

/* ------------------------------------------------------------------------------------------
 * Spatial average (replaces analysis/spatial_average.py:31-63, processframe: FastNS pairs, then
 * np.mean over each particle's neighbours and itself -- SURVEY.md section 8f row 1).
 *
 * dsg_wrap_positions: out = pos folded into [0, box) per frame (d_box (n_frames,3) float64 on the
 *   device), what FastNS does before binning; feed the result to dsg_neigh_rows with
 *   respect_pbc = 1 | 4 (minimum image, inclusive cutoff).
 * dsg_spatial_average_rows: out[i, f, c] = (values[i, f, c] + sum_{j in row(f, i)} values[j, f, c])
 *   / (counts[f, i] + 1); values / out float64 with element strides v_stride_i / v_stride_f
 *   (components contiguous, n_comp = 1 for a scalar descriptor); expected_row_len (0 = unknown)
 *   picks how many lanes share one (particle, frame) of a scalar descriptor.
 * ------------------------------------------------------------------------------------------ */
int dsg_wrap_positions(const void* d_pos, int pos_dtype, int n_frames, int n_atoms,
                       const double* d_box, void* d_out, void* stream);

int dsg_spatial_average_rows(const int32_t* d_counts, const int64_t* d_row_start,
                             const int32_t* d_rows, int n_frames, int n_cent, int n_comp,
                             const double* d_values, int64_t v_stride_i, int64_t v_stride_f,
                             double* d_out, int64_t o_stride_i, int64_t o_stride_f,
                             int expected_row_len, void* stream);

/* ------------------------------------------------------------------------------------------
 * Self time-correlation function (replaces analysis/time_correlations.py:60-92 -- SURVEY.md
 * section 8f row 4).  d_data (n_part, n_frames) float64 with row stride stride_i (time contiguous).
 * For every lag t' < max_delay (1 <= max_delay <= n_frames):
 *   c_i(t') = sum_{t < n_frames - t'} x_i(t) x_i(t + t'),  x_i = data_i - mean_t(data_i)
 *   mean_out[t'] = mean_i c_i(t'),  std_out[t'] = sqrt(mean_i (c_i(t') - mean_out[t'])^2).
 * The caller divides by (n_frames - t') and normalises by lag 0 (time_correlations.py:78-90).
 * ------------------------------------------------------------------------------------------ */
int dsg_tcf_workspace_bytes(int n_part, int n_frames, int max_delay, size_t* bytes_out);

int dsg_self_time_correlation(const double* d_data, int64_t stride_i, int n_part, int n_frames,
                              int max_delay, void* d_workspace, size_t workspace_bytes,
                              double* d_mean_out, double* d_std_out, void* stream);

/* Cross time-correlation (analysis/time_correlations.py:155-176): the same sums over all ordered
 * pairs i != j, c_ij(t') = sum_t x_i(t) x_j(t + t'); mean / std over the n_part (n_part - 1) pairs
 * per lag (n_part >= 2).  The caller divides by (n_frames - t') (no lag-0 normalisation there). */
int dsg_cross_tcf_workspace_bytes(int n_part, int n_frames, int max_delay, size_t* bytes_out);

int dsg_cross_time_correlation(const double* d_data, int64_t stride_i, int n_part, int n_frames,
                               int max_delay, void* d_workspace, size_t workspace_bytes,
                               double* d_mean_out, double* d_std_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * XTC ingest (host code; replaces the MDAnalysis XTCReader behind Trj.init_from_xtc,
 * trajectory/trajectory.py:68-79 -- SURVEY.md section 8f row 3).  Handles uncompressed
 * (natoms <= 9) and compressed frames.
 *
 * dsg_xtc_index: scans the frame headers; writes the number of frames and atoms and, if
 *   offsets_out != NULL, the byte offset of the first max_frames frames.
 * dsg_xtc_read: decodes the frames at the given offsets with n_threads host threads
 *   (<= 0: all cores).  pos_out (n_frames, n_atoms, 3) float32 in angstrom (nm * 10 in float32,
 *   as MDAnalysis converts), box_out (n_frames, 9) float32 box vectors in angstrom or NULL,
 *   step_out / time_out (n_frames) or NULL.
 * ------------------------------------------------------------------------------------------ */
int dsg_xtc_index(const char* path, int64_t* n_frames_out, int32_t* n_atoms_out,
                  int64_t* offsets_out, int64_t max_frames);

int dsg_xtc_read(const char* path, const int64_t* offsets, int64_t n_frames, int32_t n_atoms,
                 float* pos_out, float* box_out, int32_t* step_out, float* time_out,
                 int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* DYNSIGHT_B200_H */
