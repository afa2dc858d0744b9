/* zksaas_gpu.h -- C ABI of libzksaas_gpu.so: the B200 (sm_100a) replacement for the data-parallel
 * prover core of zk-SaaS (tangle-network/zk-SaaS).  The reference has no FFI of its own; each
 * entry point below replaces the *body* of one Rust function (or the arkworks call inside it)
 * and cites it as  <file>:<line>  relative to the reference root.  INTEGRATION.md shows the
 * Rust `-sys` binding and the patched function bodies.
 *
 * Conventions
 *   - Memory images are arkworks 0.4 in-memory forms, little-endian:
 *       Fr / Fq   : 4 x u64 Montgomery limbs (value * 2^256 mod p), fully reduced      32 B
 *       G1 affine : { x: Fq, y: Fq, infinity: bool }  (x@0, y@32, infinity@64)          72 B
 *       G2 affine : { x: Fq2{c0,c1}, y: Fq2{c0,c1}, infinity: bool } (infinity@128)    136 B
 *       G1 / G2 projective result: Jacobian (X, Y, Z) of 3 x 32 B / 3 x 64 B, always returned
 *       normalised (Z = 1, or (1,1,0) for the identity): a valid `Projective` whose affine form
 *       is bit-identical to arkworks' `into_affine()` of the reference result.
 *   - Host-pointer entry points (no suffix) are blocking, re-entrant and thread-safe; the caller
 *     owns every buffer and the library never retains a pointer after returning.  They use
 *     pageable or pinned host memory alike (pinned memory avoids a staging copy).
 *   - `_dev` entry points take DEVICE pointers plus a context created with zkg_ctx_create(); they
 *     enqueue work on the context's stream and return without synchronising.
 *   - Every function returns 0 (ZKG_OK) or a negative ZKG_ERR_* code; zkg_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     ZKG_ERR_CUDA.
 */
#ifndef ZKSAAS_GPU_H
#define ZKSAAS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKG_OK 0
#define ZKG_ERR_LEN_MISMATCH (-1) /* bases.len() != scalars.len(): Rust shim returns Err(min(len)) like G::msm */
#define ZKG_ERR_BAD_ARG (-2)
#define ZKG_ERR_CUDA (-3)
#define ZKG_ERR_OOM (-4)
#define ZKG_ERR_UNSUPPORTED (-5)
#define ZKG_ERR_NCCL (-6)       /* a multi-GPU entry point could not set up or run its exchange (no peer access and no NCCL) */

#define ZKG_G1_AFFINE_BYTES 72u
#define ZKG_G2_AFFINE_BYTES 136u

typedef struct zkg_ctx zkg_ctx;

/* ---- library / context management ------------------------------------------------------- */
int32_t zkg_version(void);
int32_t zkg_device_count(int32_t *count);
const char *zkg_last_error(void);
/* Context = device ordinal + stream + grow-only device workspace.  `stream` may be an existing
 * cudaStream_t (e.g. torch's current stream; pass cudaStreamLegacy = (void*)1 for the default
 * stream) or NULL to let the library create its own non-blocking stream -- in which case the caller
 * must order its own work against zkg_ctx_sync(). */
int32_t zkg_ctx_create(int32_t device, void *stream, zkg_ctx **out);
int32_t zkg_ctx_destroy(zkg_ctx *ctx);
int32_t zkg_ctx_sync(zkg_ctx *ctx);
void *zkg_ctx_stream(zkg_ctx *ctx);
/* Instrumentation for benchmarks: number of kernels this context has launched so far; and, when
 * profiling is enabled, the device time between phase boundaries of the LAST call on the context
 * (MSM: phase 0 = digits + counting sort, 1 = bucket accumulation, 2 = bucket reduction + final;
 * king pipeline: 0 = parameter tables, 1 = unpack + fft2 + powers, 2 = pack).  zkg_ctx_phase_ms
 * synchronises the stream. */
int32_t zkg_ctx_launch_count(zkg_ctx *ctx, uint64_t *count);
int32_t zkg_ctx_set_profiling(zkg_ctx *ctx, int32_t enable);
int32_t zkg_ctx_phase_ms(zkg_ctx *ctx, int32_t phase, float *ms);
/* Frees the pooled contexts the host-pointer entry points create lazily. */
int32_t zkg_shutdown(void);

/* ---- MSM: replaces `G::msm(bases, scalars)` at dist-primitives/src/dmsm/mod.rs:73 ---------
 * (ark-ec 0.4.2 VariableBaseMSM::msm for ark_bn254::{G1,G2}Projective; callers
 * groth16/src/prove.rs:52,106,154,209,219).  scalars: Fr Montgomery images.
 * Environment ZKG_AUTO_REGISTER=1 (opt-in): a base vector seen again behind the same pointer is registered transparently; every
 * later call still ships the bases, which are compared byte for byte on the device with the registered copy (a mismatch falls
 * back to the ordinary path), while the MSM runs against the prepared table -- see zkg_bases_register below. */
int32_t zkg_msm_bn254_g1(int32_t device, const void *bases, size_t base_stride, size_t n_bases,
                         const uint64_t *scalars, size_t n_scalars, uint64_t out_xyz[12]);
int32_t zkg_msm_bn254_g2(int32_t device, const void *bases, size_t base_stride, size_t n_bases,
                         const uint64_t *scalars, size_t n_scalars, uint64_t out_xyz[24]);

/* Device-resident CRS shares (the bases of groth16/src/proving_key.rs:15-45 are static across
 * proofs).  Registration uploads the bases once and PREPARES them: the table holds every window
 * shift 2^(c*w) * P_i (W = 254/c + 1 affine copies per base, e.g. 13 x 256 MiB for 2^22 G1 points --
 * sized for the B200's 180 GB).  MSMs against the handle then need a single bucket set and no
 * Horner over windows: ~1.4x faster at 2^22 points and ~2x at 2^19.  group: 1 = G1, 2 = G2. */
int32_t zkg_bases_register(int32_t device, int32_t group, const void *bases, size_t base_stride, size_t n,
                           uint64_t *handle);
/* same, from packed device bases (zkg_pack_bases_dev / zkg_fixed_base_dev output); asynchronous on ctx */
int32_t zkg_bases_register_dev(zkg_ctx *ctx, int32_t group, const void *d_bases_packed, size_t n, uint64_t *handle);
int32_t zkg_bases_release(uint64_t handle);
int32_t zkg_msm_bn254_registered(uint64_t handle, const uint64_t *scalars, size_t n_scalars, uint64_t *out_xyz);
/* device scalars / device result; partial != 0 leaves the XYZZ partial sum (see zkg_msm_combine_dev) */
int32_t zkg_msm_bn254_registered_dev(zkg_ctx *ctx, uint64_t handle, const uint64_t *d_scalars, size_t n_scalars,
                                     uint64_t *d_out, int32_t partial);

/* CRS share pre-processing: `pp.det_pack::<G>(chunk)` for every l-chunk of a proving-key
 * query, then the per-party affine shares -- groth16/src/proving_key.rs:72-104 (pack_from_arkworks_proving_key),
 * secret-sharing/src/pss.rs:69-87.  bases: n = chunks*l arkworks Affine images; out_by_party[i] (i < 4l):
 * `chunks` Affine images, i.e. party i's `s` / `u` / `w` / `h` (G1) or `v` (G2) vector. */
int32_t zkg_crs_det_pack_bn254(int32_t device, int32_t group, const void *bases, size_t base_stride, size_t n,
                               uint32_t l, void *const *out_by_party, size_t out_stride);

/* Device-pointer MSM.  d_bases: packed affine (x,y) Montgomery, 64 B (G1) / 128 B (G2) per point,
 * infinity encoded as (0,0) -- produce it with zkg_pack_bases_dev.  d_scalars: n x 32 B.
 * d_out_xyz: 96 B / 192 B on the device. */
int32_t zkg_pack_bases_dev(zkg_ctx *ctx, int32_t group, const void *d_bases_ark, size_t base_stride, size_t n,
                           void *d_bases_packed);
int32_t zkg_msm_bn254_g1_dev(zkg_ctx *ctx, const void *d_bases_packed, const uint64_t *d_scalars, size_t n,
                             uint64_t *d_out_xyz);
int32_t zkg_msm_bn254_g2_dev(zkg_ctx *ctx, const void *d_bases_packed, const uint64_t *d_scalars, size_t n,
                             uint64_t *d_out_xyz);
/* Partial MSM for multi-GPU sharding: same as above but leaves the (un-normalised) XYZZ partial sum
 * (4 x 32 B / 4 x 64 B) in d_out_xyzz; partials from all ranks are combined with
 * zkg_msm_combine_dev after an all-gather. */
int32_t zkg_msm_bn254_partial_dev(zkg_ctx *ctx, int32_t group, const void *d_bases_packed, const uint64_t *d_scalars,
                                  size_t n, uint64_t *d_out_xyzz);
int32_t zkg_msm_combine_dev(zkg_ctx *ctx, int32_t group, const uint64_t *d_partials_xyzz, size_t n_partials,
                            uint64_t *d_out_xyz);
/* bases[i] = scalars[i] * generator, written in packed device form (synthetic CRS for benches
 * and tests; also the kernel behind `gen * x` in MsmMask::sample, dmsm/mod.rs:32). */
int32_t zkg_fixed_base_dev(zkg_ctx *ctx, int32_t group, const uint64_t *d_scalars, size_t n, void *d_bases_packed);

/* ---- client-side FFT step: replaces fft1_in_place, dist-primitives/src/dfft/mod.rs:178-208 ----
 * px: this party's share vector (mbyl = m/l elements), transformed in place.
 * pre_scale (nullable): every element is first multiplied by it (d_ifft's size_inv, dfft/mod.rs:159).
 * in_mask (nullable, mbyl elements): added after the transform (fft2_with_rearrange's masking,
 * dfft/mod.rs:254-258).  gen: dom.group_gen() or dom.group_gen_inv(). */
int32_t zkg_fft1_bn254(int32_t device, uint64_t *px, size_t mbyl, uint32_t l, const uint64_t gen[4],
                       const uint64_t *pre_scale, const uint64_t *in_mask);
int32_t zkg_fft1_bn254_dev(zkg_ctx *ctx, uint64_t *d_px, size_t mbyl, uint32_t l, const uint64_t gen[4],
                           const uint64_t *pre_scale, const uint64_t *d_in_mask);

/* fft1 of ONE lane sharded over the GPUs of a box (north_star "four-step with an NCCL
 * all-to-all transpose").  Rank g holds the contiguous block px[g*block_len, (g+1)*block_len) of the
 * lane (block_len = (m/l)/n_ranks).  shard_local transforms the block in place into the send
 * buffer (inner size-block_len transform + twiddles; chunk d of it goes to rank d); after the
 * all-to-all, shard_outer turns the n_ranks x cols receive buffer (chunk g from rank g) into
 * out[k1*cols + j] = X[(first_col + j) + block_len*k1], where fft1_in_place(px)[k] = X[(k+1) mod (m/l)]
 * (dfft/mod.rs:178-208) and first_col = rank * cols.  All sizes are powers of two. */
int32_t zkg_fft1_shard_local_bn254_dev(zkg_ctx *ctx, uint64_t *d_block, size_t block_len, uint32_t l, uint32_t n_ranks,
                                       uint32_t rank, const uint64_t gen[4], const uint64_t *pre_scale);
int32_t zkg_fft1_shard_outer_bn254_dev(zkg_ctx *ctx, const uint64_t *d_recv, size_t cols, size_t block_len, uint32_t l,
                                       uint32_t n_ranks, const uint64_t gen[4], uint64_t *d_out);

/* ---- king closure of fft2_with_rearrange: dist-primitives/src/dfft/mod.rs:264-304 -----------
 * shares_by_party[r]: the mbyl-element vector received from party parties[r] (r < n_recv; n_recv
 * == n = 4l uses unpack2, fewer uses the Lagrange matrix of pss.rs:170-207).  rand: mbyl x t
 * random field elements for re-packing (column i uses rand[i*t .. i*t+t)); the host RNG stays
 * in Rust.  out_by_party[p]: mbyl elements for each of the n parties.  g: coset shift (one() for
 * d_fft).  rearrange != 0 selects the bit-reversed, strided re-packing (dfft/mod.rs:284-300). */
int32_t zkg_king_fft2_bn254(int32_t device, const uint64_t *const *shares_by_party, const uint32_t *parties,
                            uint32_t n_recv, size_t mbyl, uint32_t l, const uint64_t gen[4], const uint64_t g[4],
                            int32_t rearrange, const uint64_t *rand, uint64_t *const *out_by_party);
/* Device form: d_shares is party-major contiguous (n_recv x mbyl), d_out is party-major (n x mbyl). */
int32_t zkg_king_fft2_bn254_dev(zkg_ctx *ctx, const uint64_t *d_shares, const uint32_t *parties, uint32_t n_recv,
                                size_t mbyl, uint32_t l, const uint64_t gen[4], const uint64_t g[4],
                                int32_t rearrange, const uint64_t *d_rand, uint64_t *d_out);

/* Column-range halves of the same pipeline, for sharding ONE king over the GPUs of a box.
 * Stage 1 takes the share columns [col0, col0+cols) (d_shares_local: n_recv x cols, party-major) and
 * scatters unpack -> fft2 -> g^i into the FULL-size buffer d_S_full (m = mbyl*l elements in pack order,
 * zeroed by the caller): every slot is written by exactly one rank, so a sum reduce-scatter over the
 * ranks (one NCCL collective) hands each rank the contiguous slice of its own output columns.  Stage 2
 * packs `cols` output columns: d_S_local (cols*l), d_rand_local (cols*t) -> d_out_local (n x cols). */
int32_t zkg_king_stage1_bn254_dev(zkg_ctx *ctx, const uint64_t *d_shares_local, const uint32_t *parties,
                                  uint32_t n_recv, size_t col0, size_t cols, size_t mbyl, uint32_t l,
                                  const uint64_t gen[4], const uint64_t g[4], int32_t rearrange, uint64_t *d_S_full);
int32_t zkg_king_stage2_bn254_dev(zkg_ctx *ctx, const uint64_t *d_S_local, const uint64_t *d_rand_local, size_t cols,
                                  uint32_t l, uint64_t *d_out_local);

/* ---- king closure of deg_red: dist-primitives/src/utils/deg_red.rs:103-111 ------------------ */
int32_t zkg_deg_red_king_bn254(int32_t device, const uint64_t *const *shares_by_party, const uint32_t *parties,
                               uint32_t n_recv, size_t cols, uint32_t l, const uint64_t *rand,
                               uint64_t *const *out_by_party);
int32_t zkg_deg_red_king_bn254_dev(zkg_ctx *ctx, const uint64_t *d_shares, const uint32_t *parties, uint32_t n_recv,
                                   size_t cols, uint32_t l, const uint64_t *d_rand, uint64_t *d_out);

/* ---- king closure of d_pp (SURVEY.md 8f "next", rank 1): dist-primitives/src/dpp/mod.rs:41-76 ----
 * shares_by_party[r]: 2*cols elements (the party's num shares followed by its den shares).  The king
 * unpacks both halves, divides (batch inversion), takes the running product over all cols*l secrets
 * and re-packs with the supplied randomness.  A zero denominator returns ZKG_ERR_BAD_ARG (the
 * reference's `.inverse().unwrap()` panics). */
int32_t zkg_dpp_king_bn254(int32_t device, const uint64_t *const *shares_by_party, const uint32_t *parties,
                           uint32_t n_recv, size_t cols, uint32_t l, const uint64_t *rand,
                           uint64_t *const *out_by_party);

/* Offline packing of whole vectors into the n parties' share vectors.
 * layout 0: chunks of l consecutive values, last chunk zero-padded -- pack_from_witness
 *           (groth16/examples/sha256.rs:131-156) and pack_vec + transpose (utils/pack.rs:8-20);
 * layout 1: bit-reverse x (len a power of two), column i packs (x'[i], x'[i + len/l], ...) -- the
 *           `pack` closure of QAP::pss (groth16/src/qap.rs:99-112).
 * rand: ceil(len/l) * t host-RNG points; out_by_party[p] (p < 4l): ceil(len/l) shares of party p. */
int32_t zkg_pss_pack_vec_bn254_fr(int32_t device, uint32_t l, int32_t layout, const uint64_t *x, size_t len,
                                  const uint64_t *rand, uint64_t *const *out_by_party);

/* ---- PackedSharingParams transforms over Fr, batched over `cols` columns -------------------
 * secret-sharing/src/pss.rs: pack :90-122 (rand != NULL), det_pack :69-87 (rand == NULL),
 * unpack :125-138, unpack2 :141-166.  Column-major contiguous: column c reads secrets[c*l..],
 * rand[c*t..], writes shares[c*n..]. */
int32_t zkg_pss_pack_bn254_fr(int32_t device, uint32_t l, const uint64_t *secrets, const uint64_t *rand,
                              uint64_t *shares, size_t cols);
int32_t zkg_pss_unpack_bn254_fr(int32_t device, uint32_t l, const uint64_t *shares, uint64_t *secrets, size_t cols);
int32_t zkg_pss_unpack2_bn254_fr(int32_t device, uint32_t l, const uint64_t *shares, uint64_t *secrets, size_t cols);

/* ---- stand-alone pieces (FftMask::sample dfft/mod.rs:30-85, QAP::pss groth16/src/qap.rs:101) ---- */
/* fft2_in_place, dfft/mod.rs:210-237 (s1: m elements, in place) */
int32_t zkg_fft2_bn254(int32_t device, uint64_t *s1, size_t m, uint32_t l, const uint64_t gen[4]);
/* Radix2EvaluationDomain::distribute_powers (dfft/mod.rs:49,279): v[i] *= g^i */
int32_t zkg_distribute_powers_bn254(int32_t device, uint64_t *v, size_t n, const uint64_t g[4]);
/* fft_in_place_rearrange, dfft/mod.rs:322-335 (bit-reversal permutation, n = 2^k) */
int32_t zkg_bitrev_bn254(int32_t device, uint64_t *v, size_t n);
/* Radix2EvaluationDomain::{fft,ifft}_in_place on n = 2^k elements, optional coset offset (NULL = 1);
 * the plain transforms of groth16/src/ext_wit.rs:204-285 and the reconstruction side of the tests. */
int32_t zkg_fr_fft_bn254(int32_t device, uint64_t *v, size_t n, const uint64_t *offset, int32_t inverse);

/* ---- wire format: ark-serialize compressed Fr = 32 LE bytes of the canonical
 * value (the payload of mpc-net/src/ser_net.rs:25,40,112,119 frames after the u64 length prefix).
 * from_wire fails with ZKG_ERR_BAD_ARG if an element is >= r, as arkworks' deserializer does. */
int32_t zkg_fr_from_wire_bn254(int32_t device, const void *wire, uint64_t *out_mont, size_t n);
int32_t zkg_fr_to_wire_bn254(int32_t device, const uint64_t *in_mont, void *wire, size_t n);


/* ---- several GPUs of one box behind ONE call (SURVEY.md 8b last row, 8e) ------------------------------------------
 * The Rust callers (`d_msm` at dist-primitives/src/dmsm/mod.rs:73, the king closure at dfft/mod.rs:264-304, the client
 * step at :121/:162) make one call per operation from one process; these entry points take a device LIST instead of a
 * device ordinal and shard that one operation.  Same argument meaning, same results bit for bit as the single-GPU
 * entry points above (tests/test_gpu_multi.py).  n_devices must be a power of two for the king / fft1 forms.
 *   MSM        point-range split; every GPU runs the whole pipeline on its slice (its slice of the host buffers crosses
 *              its own PCIe link, driven by its own host thread); the 128 / 256-byte XYZZ partials go to devices[0] over
 *              NVLink peer copies and are added there.  No bucket exchange: MSM is linear.
 *   king       share-column split; stage 1 (unpack, fft2, g^i) stores every value straight into the memory of the GPU
 *              that owns its output column -- NVLink peer stores ARE the rotate / bit-reverse / stride permutation of
 *              dfft/mod.rs:284-300, so there is no separate collective -- then every GPU packs its own columns.
 *   fft1       four-step over one lane: inner transforms on contiguous blocks, twiddle multiplication fused with the
 *              peer-store all-to-all, G-point outer transforms; the result is laid back into px in fft1 order.
 * Peer access between the listed GPUs is required (NVLink / NVSwitch on a B200 box); without it: ZKG_ERR_NCCL. */
int32_t zkg_msm_bn254_g1_sharded(const int32_t *devices, int32_t n_devices, const void *bases, size_t base_stride,
                                 size_t n_bases, const uint64_t *scalars, size_t n_scalars, uint64_t out_xyz[12]);
int32_t zkg_msm_bn254_g2_sharded(const int32_t *devices, int32_t n_devices, const void *bases, size_t base_stride,
                                 size_t n_bases, const uint64_t *scalars, size_t n_scalars, uint64_t out_xyz[24]);
/* Registers consecutive point ranges of one CRS share on the listed GPUs; the handle works with zkg_msm_bn254_registered
 * (host scalars; each GPU receives only its range) and zkg_bases_release.  Not usable with the `_dev` form. */
int32_t zkg_bases_register_sharded(const int32_t *devices, int32_t n_devices, int32_t group, const void *bases,
                                   size_t base_stride, size_t n, uint64_t *handle);
int32_t zkg_king_fft2_bn254_sharded(const int32_t *devices, int32_t n_devices, const uint64_t *const *shares_by_party,
                                    const uint32_t *parties, uint32_t n_recv, size_t mbyl, uint32_t l,
                                    const uint64_t gen[4], const uint64_t g[4], int32_t rearrange, const uint64_t *rand,
                                    uint64_t *const *out_by_party);
int32_t zkg_deg_red_king_bn254_sharded(const int32_t *devices, int32_t n_devices, const uint64_t *const *shares_by_party,
                                       const uint32_t *parties, uint32_t n_recv, size_t cols, uint32_t l,
                                       const uint64_t *rand, uint64_t *const *out_by_party);
int32_t zkg_fft1_bn254_sharded(const int32_t *devices, int32_t n_devices, uint64_t *px, size_t mbyl, uint32_t l,
                               const uint64_t gen[4], const uint64_t *pre_scale, const uint64_t *in_mask);

/* The same exchanges for ONE PROCESS PER GPU (torchrun-style launchers): buffers that the kernels of the other ranks
 * store into are allocated with zkg_shared_alloc, their 64-byte CUDA IPC handles are exchanged by the host (any
 * transport), and each rank maps its peers' buffers with zkg_shared_open.  The caller orders "all ranks have run the
 * scatter step" before the consuming step (one barrier; e.g. a 4-byte NCCL all-reduce on the same stream).
 *   zkg_king_stage1_scatter_bn254_dev: like zkg_king_stage1_bn254_dev, but d_S_by_rank[r] is rank r's OWN segment of the
 *   pack-order buffer (m / n_ranks elements: its output columns), this rank's entry being its local allocation.
 *   zkg_fft1_shard_local_scatter_bn254_dev: zkg_fft1_shard_local_bn254_dev fused with the all-to-all: chunk d of the
 *   twiddled inner transform is stored into d_recv_by_rank[d] at chunk position `rank` (block_len elements per buffer). */
int32_t zkg_shared_alloc(zkg_ctx *ctx, size_t bytes, void **d_ptr, uint8_t ipc_handle[64]);
int32_t zkg_shared_open(zkg_ctx *ctx, const uint8_t ipc_handle[64], void **d_ptr);
int32_t zkg_shared_close(zkg_ctx *ctx, void *d_ptr);
int32_t zkg_shared_free(zkg_ctx *ctx, void *d_ptr);
int32_t zkg_king_stage1_scatter_bn254_dev(zkg_ctx *ctx, const uint64_t *d_shares_local, const uint32_t *parties,
                                          uint32_t n_recv, size_t col0, size_t cols, size_t mbyl, uint32_t l,
                                          const uint64_t gen[4], const uint64_t g[4], int32_t rearrange,
                                          void *const *d_S_by_rank, uint32_t n_ranks);
int32_t zkg_fft1_shard_local_scatter_bn254_dev(zkg_ctx *ctx, uint64_t *d_block, size_t block_len, uint32_t l,
                                               uint32_t n_ranks, uint32_t rank, const uint64_t gen[4],
                                               const uint64_t *pre_scale, void *const *d_recv_by_rank);

/* ---- king side of d_msm: dist-primitives/src/dmsm/mod.rs:85-87 ---------------------------------
 * `pp.unpack_missing_shares(&rs.shares, &rs.parties)` over GROUP elements (secret-sharing/src/pss.rs:141-166 when all
 * n = 4l shares arrived, the Lagrange path :170-221 otherwise -- d_msm tolerates dropouts like d_fft) followed by
 * `result.iter().sum()`.  Also the A / B / C recombination of groth16/examples/sha256.rs:375-377.
 * shares_xyz: n_recv Projective images (Jacobian X, Y, Z of 3 x 32 B / 3 x 64 B; any Z, identity Z = 0) in the order of
 * `parties` (NULL = parties 0..n-1).  out_unpacked_xyz (nullable): the l unpacked points; out_sum_xyz (nullable): their
 * sum (the value the king replicates to every party, :87).  Outputs are normalised. */
int32_t zkg_pss_unpack2_bn254_g1(int32_t device, uint32_t l, const uint64_t *shares_xyz, const uint32_t *parties,
                                 uint32_t n_recv, uint64_t *out_unpacked_xyz, uint64_t *out_sum_xyz);
int32_t zkg_pss_unpack2_bn254_g2(int32_t device, uint32_t l, const uint64_t *shares_xyz, const uint32_t *parties,
                                 uint32_t n_recv, uint64_t *out_unpacked_xyz, uint64_t *out_sum_xyz);

/* ---- wire format of group elements: ark-serialize 0.4 COMPRESSED points, the payload d_msm
 * ships through mpc-net/src/ser_net.rs:25 (serialize_compressed) and reads back at :40 / :119 (deserialize_compressed =
 * Compress::Yes + Validate::Yes); call sites dist-primitives/src/dmsm/mod.rs:79-81, :90-92.
 *   G1: 32 B = x canonical little-endian;  G2: 64 B = x.c0 then x.c1;  flags in the two top bits of the LAST byte:
 *   0x80 = y > -y ("negative"; Fq2 compares c1 first), 0x40 = point at infinity (x = 0).
 * to_wire takes Projective images (any Z); from_wire returns normalised images and fails with ZKG_ERR_BAD_ARG on an
 * invalid encoding, as arkworks does: both flags set, x >= q, x not on the curve, (G2) point outside the r-torsion. */
int32_t zkg_g1_to_wire_bn254(int32_t device, const uint64_t *points_xyz, void *wire, size_t n);
int32_t zkg_g1_from_wire_bn254(int32_t device, const void *wire, uint64_t *points_xyz, size_t n);
int32_t zkg_g2_to_wire_bn254(int32_t device, const uint64_t *points_xyz, void *wire, size_t n);
int32_t zkg_g2_from_wire_bn254(int32_t device, const void *wire, uint64_t *points_xyz, size_t n);

/* ---- share-wise h = a*b - c of the QAP, fused with the out-mask additions of the three d_fft calls before it
 * (SURVEY.md 8f row 1): groth16/src/ext_wit.rs:173-177 (circom_h), :82-86 (libsnark_h, with factor = 1/Z(g));
 * the mask adds are dist-primitives/src/dfft/mod.rs:313-317.
 *   out[i] = ((a[i] + mask_a[i]) * (b[i] + mask_b[i]) - (c[i] + mask_c[i])) * factor
 * mask_* and factor (one Fr image, host memory in both forms) are nullable; out may alias an input. */
int32_t zkg_qap_h_bn254(int32_t device, const uint64_t *a, const uint64_t *b, const uint64_t *c, const uint64_t *mask_a,
                        const uint64_t *mask_b, const uint64_t *mask_c, const uint64_t *factor, uint64_t *out, size_t n);
int32_t zkg_qap_h_bn254_dev(zkg_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, const uint64_t *d_c,
                            const uint64_t *d_mask_a, const uint64_t *d_mask_b, const uint64_t *d_mask_c,
                            const uint64_t *factor, uint64_t *d_out, size_t n);

/* ---- offline masks on the device, random draws supplied by the caller's RNG:
 * FftMask::sample, dist-primitives/src/dfft/mod.rs:30-85: mask_values (m), rand_in / rand_out (m/l * t packing draws) ->
 * the n parties' in_mask / out_mask vectors (m/l each).  One upload, one download; pack, fft2, powers, negation,
 * bit-reversal and the strided re-pack never leave the GPU.
 * DegRedMask::sample over Fr with gen = 1, dist-primitives/src/utils/deg_red.rs:40-66: mask_values (num*l). */
int32_t zkg_fft_mask_sample_bn254(int32_t device, int32_t rearrange, const uint64_t g[4], const uint64_t gen[4], size_t m,
                                  uint32_t l, const uint64_t *mask_values, const uint64_t *rand_in, const uint64_t *rand_out,
                                  uint64_t *const *in_by_party, uint64_t *const *out_by_party);
int32_t zkg_deg_red_mask_sample_bn254(int32_t device, size_t num, uint32_t l, const uint64_t *mask_values,
                                      const uint64_t *rand_in, const uint64_t *rand_out, uint64_t *const *in_by_party,
                                      uint64_t *const *out_by_party);

/* ---- element-wise helpers used by kernel unit tests (out[i] = a[i] op b[i]; op: 0 mul, 1 add,
 * 2 sub; unit-test views of the routines the kernels are built from: 3 dedicated squaring a^2, 4 the two-term
 * inner product a*b - (a+1)*b (= -b, one Montgomery reduction), 5 inverse of a (0 -> 0); field: 0 Fr, 1 Fq) ---- */
int32_t zkg_field_op(int32_t device, int32_t field, int32_t op, const uint64_t *a, const uint64_t *b, uint64_t *out,
                     size_t n);
/* device form; also the share-wise h = a*b - c of groth16/src/ext_wit.rs:82-86,173-177 and the mask adds */
int32_t zkg_field_op_dev(zkg_ctx *ctx, int32_t field, int32_t op, const uint64_t *d_a, const uint64_t *d_b,
                         uint64_t *d_out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* ZKSAAS_GPU_H */
