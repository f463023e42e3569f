// Header-only C++ host mirror of the reference-facing API for the MSM hot path, over the C ABI in
// include/zkmsm.h.  Names follow curve25519-dalek's public items for this path (CompressedRistretto,
// Scalar, VartimeMultiscalarMul::{vartime_multiscalar_mul, optional_multiscalar_mul}) as recalled from
// public knowledge -- the upstream source is not mounted, so there is no file:line
// to cite.  Nothing is computed on the CPU here; every call forwards to libzkmsm.so.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/zkmsm.h"

namespace zkvm_b200 {

using Scalar = std::array<uint8_t, 32>;               // little-endian, taken modulo the group order
using CompressedRistretto = std::array<uint8_t, 32>;  // RFC 9496 encoding

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

class Context {
  public:
    explicit Context(int device = 0) {
        int rc = zk_ctx_create(device, &h_);
        if (rc != ZK_OK) throw Error(rc, std::string("zk_ctx_create: ") + zk_status_str(rc) + " (no CPU fallback)");
    }
    ~Context() { zk_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    zk_ctx* raw() const { return h_; }
    void check(int rc) const {
        if (rc != ZK_OK) throw Error(rc, std::string(zk_status_str(rc)) + ": " + zk_last_error(h_));
    }
  private:
    zk_ctx* h_ = nullptr;
};

// Device-resident Vec<RistrettoPoint>.
class PointTable {
  public:
    PointTable(Context& ctx, size_t capacity = 0) : ctx_(ctx) { ctx_.check(zk_table_create(ctx.raw(), capacity, &h_)); }
    ~PointTable() { zk_table_destroy(h_); }
    PointTable(const PointTable&) = delete;
    PointTable& operator=(const PointTable&) = delete;
    size_t size() const { return zk_table_len(h_); }
    // Returns the index of the first invalid encoding, or nullopt when all n were appended.
    std::optional<size_t> append_compressed(const CompressedRistretto* pts, size_t n) {
        size_t bad = 0;
        int rc = zk_table_append_compressed(ctx_.raw(), h_, reinterpret_cast<const uint8_t*>(pts), n, &bad);
        if (rc == ZK_ERR_INVALID_POINT) return bad;
        ctx_.check(rc);
        return std::nullopt;
    }
    void append_uniform(const uint8_t* bytes64, size_t n) { ctx_.check(zk_table_append_uniform(ctx_.raw(), h_, bytes64, n)); }
    // Already-decompressed points as X, Y, Z, T (4 x 32-byte canonical field elements each); index of the first bad one.
    std::optional<size_t> append_extended(const uint8_t* ext128, size_t n) {
        size_t bad = 0;
        int rc = zk_table_append_extended(ctx_.raw(), h_, ext128, n, &bad);
        if (rc == ZK_ERR_INVALID_POINT) return bad;
        ctx_.check(rc);
        return std::nullopt;
    }
    // Same input for points known to be valid representatives (values of a RistrettoPoint): only Z = 0 is rejected.
    std::optional<size_t> append_extended_unchecked(const uint8_t* ext128, size_t n) {
        size_t bad = 0;
        int rc = zk_table_append_extended_unchecked(ctx_.raw(), h_, ext128, n, &bad);
        if (rc == ZK_ERR_INVALID_POINT) return bad;
        ctx_.check(rc);
        return std::nullopt;
    }
    // Window expansion for static generator sets (the device-side VartimePrecomputedMultiscalarMul); c = 0 picks the width.
    void precompute(int c = 0) { ctx_.check(zk_table_precompute(ctx_.raw(), h_, c)); }
    int precomputed_window() const { return zk_table_precomputed_window(h_); }
    std::vector<CompressedRistretto> compress(size_t offset, size_t n) const {
        std::vector<CompressedRistretto> out(n);
        ctx_.check(zk_table_compress(ctx_.raw(), h_, offset, n, reinterpret_cast<uint8_t*>(out.data())));
        return out;
    }
    const zk_table* raw() const { return h_; }
  private:
    Context& ctx_;
    zk_table* h_ = nullptr;
};

struct RistrettoPoint {
    // sum scalars[i] * table[offset + i]
    static CompressedRistretto vartime_multiscalar_mul(Context& ctx, const std::vector<Scalar>& scalars, const PointTable& points,
                                                       size_t offset = 0) {
        CompressedRistretto out{};
        ctx.check(zk_msm_vartime_table(ctx.raw(), reinterpret_cast<const uint8_t*>(scalars.data()), points.raw(), offset,
                                       scalars.size(), out.data()));
        return out;
    }
    // sum scalars[i] * decompress(points[i]); nullopt if any encoding is invalid
    static std::optional<CompressedRistretto> optional_multiscalar_mul(Context& ctx, const std::vector<Scalar>& scalars,
                                                                       const std::vector<CompressedRistretto>& points) {
        if (scalars.size() != points.size()) throw Error(ZK_ERR_ARG, "scalars/points length mismatch");
        CompressedRistretto out{};
        int rc = zk_msm_vartime(ctx.raw(), reinterpret_cast<const uint8_t*>(scalars.data()),
                                reinterpret_cast<const uint8_t*>(points.data()), scalars.size(), out.data());
        if (rc == ZK_ERR_INVALID_POINT) return std::nullopt;
        ctx.check(rc);
        return out;
    }
    // Cached generators first, then the proof's own compressed points (bulletproofs' verification shape).
    static std::optional<CompressedRistretto> mixed_multiscalar_mul(Context& ctx, const std::vector<Scalar>& static_scalars,
                                                                    const PointTable& table, size_t offset,
                                                                    const std::vector<Scalar>& dyn_scalars,
                                                                    const std::vector<CompressedRistretto>& dyn_points) {
        if (dyn_scalars.size() != dyn_points.size()) throw Error(ZK_ERR_ARG, "dynamic scalars/points length mismatch");
        CompressedRistretto out{};
        int rc = zk_msm_vartime_mixed(ctx.raw(), reinterpret_cast<const uint8_t*>(static_scalars.data()), table.raw(), offset,
                                      static_scalars.size(), reinterpret_cast<const uint8_t*>(dyn_scalars.data()),
                                      reinterpret_cast<const uint8_t*>(dyn_points.data()), dyn_points.size(), out.data());
        if (rc == ZK_ERR_INVALID_POINT) return std::nullopt;
        ctx.check(rc);
        return out;
    }
    // m independent MSMs in one device pass; MSM k covers terms [seg[k], seg[k+1]).  nullopt where an encoding was invalid.
    static std::vector<std::optional<CompressedRistretto>> batch_optional_multiscalar_mul(
        Context& ctx, const std::vector<Scalar>& scalars, const std::vector<CompressedRistretto>& points, const std::vector<uint64_t>& seg) {
        if (seg.size() < 2 || scalars.size() != points.size() || seg.back() != scalars.size()) throw Error(ZK_ERR_ARG, "bad segments");
        size_t m = seg.size() - 1;
        std::vector<CompressedRistretto> out(m); std::vector<uint8_t> valid(m);
        int rc = zk_msm_vartime_batch(ctx.raw(), reinterpret_cast<const uint8_t*>(scalars.data()),
                                      reinterpret_cast<const uint8_t*>(points.data()), seg.data(), m,
                                      reinterpret_cast<uint8_t*>(out.data()), valid.data());
        if (rc != ZK_OK && rc != ZK_ERR_INVALID_POINT) ctx.check(rc);
        std::vector<std::optional<CompressedRistretto>> res(m);
        for (size_t k = 0; k < m; k++) if (valid[k]) res[k] = out[k];
        return res;
    }
    // m MSMs over the SAME cached points (m proofs against one generator set).
    static std::vector<CompressedRistretto> batch_vartime_multiscalar_mul(Context& ctx, const std::vector<Scalar>& scalars,
                                                                          const PointTable& table, const std::vector<uint64_t>& seg,
                                                                          size_t offset = 0) {
        if (seg.size() < 2 || seg.back() != scalars.size()) throw Error(ZK_ERR_ARG, "bad segments");
        std::vector<CompressedRistretto> out(seg.size() - 1);
        ctx.check(zk_msm_vartime_table_batch(ctx.raw(), reinterpret_cast<const uint8_t*>(scalars.data()), table.raw(), offset,
                                             seg.data(), seg.size() - 1, reinterpret_cast<uint8_t*>(out.data())));
        return out;
    }
    static bool is_identity(const CompressedRistretto& c) { return zk_encoding_is_identity(c.data()) != 0; }
    // sum of decompressed points, compressed (at most 1024); nullopt if an encoding is invalid
    static std::optional<CompressedRistretto> sum(Context& ctx, const std::vector<CompressedRistretto>& points) {
        CompressedRistretto out{};
        int rc = zk_sum_compressed(ctx.raw(), reinterpret_cast<const uint8_t*>(points.data()), points.size(), out.data());
        if (rc == ZK_ERR_INVALID_POINT) return std::nullopt;
        ctx.check(rc);
        return out;
    }
};

// Several GPUs of one box behind one call (BASELINE.json config 5): point-range shards, one gather of 128-byte partials.
class MultiGpu {
  public:
    enum class Gather { Peer = 0, Nccl = 1 };
    explicit MultiGpu(const std::vector<int>& devices, Gather gather = Gather::Peer) {
        int rc = zk_mgpu_create(devices.data(), (int)devices.size(), &h_);
        if (rc != ZK_OK) throw Error(rc, std::string("zk_mgpu_create: ") + zk_status_str(rc) + " (no CPU fallback)");
        if (gather != Gather::Peer) check(zk_mgpu_set_gather(h_, (int)gather));
    }
    ~MultiGpu() { zk_mgpu_destroy(h_); }
    MultiGpu(const MultiGpu&) = delete;
    MultiGpu& operator=(const MultiGpu&) = delete;
    zk_mgpu* raw() const { return h_; }
    int device_count() const { return zk_mgpu_device_count(h_); }
    void check(int rc) const {
        if (rc != ZK_OK) throw Error(rc, std::string(zk_status_str(rc)) + ": " + zk_mgpu_last_error(h_));
    }
    // sum scalars[i] * decompress(points[i]) over all devices; nullopt if any encoding is invalid
    std::optional<CompressedRistretto> optional_multiscalar_mul(const Scalar* scalars, const CompressedRistretto* points, size_t n) {
        CompressedRistretto out{};
        int rc = zk_mgpu_msm_vartime(h_, reinterpret_cast<const uint8_t*>(scalars), reinterpret_cast<const uint8_t*>(points), n, out.data());
        if (rc == ZK_ERR_INVALID_POINT) return std::nullopt;
        check(rc);
        return out;
    }
  private:
    zk_mgpu* h_ = nullptr;
};

// Vec<RistrettoPoint> sharded by index range over the devices of a MultiGpu.
class MultiGpuTable {
  public:
    MultiGpuTable(MultiGpu& mg, size_t capacity = 0) : mg_(mg) { mg_.check(zk_mgpu_table_create(mg.raw(), capacity, &h_)); }
    ~MultiGpuTable() { zk_mgpu_table_destroy(h_); }
    MultiGpuTable(const MultiGpuTable&) = delete;
    MultiGpuTable& operator=(const MultiGpuTable&) = delete;
    size_t size() const { return zk_mgpu_table_len(h_); }
    std::optional<size_t> append_compressed(const CompressedRistretto* pts, size_t n) {
        size_t bad = 0;
        int rc = zk_mgpu_table_append_compressed(h_, reinterpret_cast<const uint8_t*>(pts), n, &bad);
        if (rc == ZK_ERR_INVALID_POINT) return bad;
        mg_.check(rc);
        return std::nullopt;
    }
    void append_uniform(const uint8_t* bytes64, size_t n) { mg_.check(zk_mgpu_table_append_uniform(h_, bytes64, n)); }
    CompressedRistretto vartime_multiscalar_mul(const Scalar* scalars, size_t offset, size_t n) {
        CompressedRistretto out{};
        mg_.check(zk_mgpu_msm_vartime_table(mg_.raw(), reinterpret_cast<const uint8_t*>(scalars), h_, offset, n, out.data()));
        return out;
    }
    // cached generators [offset, offset + n_static) + the proof's own compressed points; nullopt if one of those is invalid
    std::optional<CompressedRistretto> mixed_multiscalar_mul(const Scalar* static_scalars, size_t offset, size_t n_static,
                                                             const Scalar* dyn_scalars, const CompressedRistretto* dyn_points, size_t n_dyn) {
        CompressedRistretto out{};
        int rc = zk_mgpu_msm_vartime_mixed(mg_.raw(), reinterpret_cast<const uint8_t*>(static_scalars), h_, offset, n_static,
                                           reinterpret_cast<const uint8_t*>(dyn_scalars), reinterpret_cast<const uint8_t*>(dyn_points), n_dyn, out.data());
        if (rc == ZK_ERR_INVALID_POINT) return std::nullopt;
        mg_.check(rc);
        return out;
    }
  private:
    MultiGpu& mg_;
    zk_mgpu_table* h_ = nullptr;
};

}  // namespace zkvm_b200
