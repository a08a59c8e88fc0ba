// Dense_Stack.h -- host-side plumbing shared by the drop-in entry points (Alignment_Demons.cc, Warp_Images_Field.cc):
// image stacks (planar_image_collection) <-> dense [z][y][x][c] volumes in pooled PINNED host memory, slice-parallel on
// the host threads, and the description of a regular stack as a dcma_b200_grid.  Header-only; not part of the public
// surface (the reference has no counterpart: it marshals through std::vector, src/Alignment_Demons.h:139-204).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <iterator>
#include <list>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "YgorImages.h"
#include "YgorMath.h"

#include "dcma_b200/demons_c_api.h"

namespace dcma_b200_host {
namespace detail {

// ---------------------------------------------------------------------------------------------------------------------
// Host-side plumbing of the drop-in: dense volumes live in PINNED host memory (dcma_b200_host_alloc; kept in a small
// process-lifetime pool, because page-locking 1.6 GB costs about as much as it saves on one transfer), and the copies
// between image stacks and dense volumes run slice-parallel on the host threads.
// ---------------------------------------------------------------------------------------------------------------------
inline unsigned worker_count() {
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(16u, hw ? hw : 4u));
}
template <class F>
void parallel_for(int64_t n, F f) {
    const unsigned nt = static_cast<unsigned>(std::min<int64_t>(worker_count(), n));
    if (nt <= 1) {
        for (int64_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            try {
                for (int64_t i = t; i < n; i += nt) f(i);
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu);
                if (!err) err = std::current_exception();
            }
        });
    for (auto &x : th) x.join();
    if (err) std::rethrow_exception(err);
}

// pinned host buffers, reused across registrations of this process
class PinnedPool {
    struct Buf {
        void *p;
        size_t bytes;
        bool busy;
    };
    std::vector<Buf> bufs_;
    std::mutex mu_;

  public:
    ~PinnedPool() {
        for (auto &b : bufs_) dcma_b200_host_free(b.p);
    }
    void *acquire(size_t bytes) {
        std::lock_guard<std::mutex> lk(mu_);
        Buf *best = nullptr;
        for (auto &b : bufs_)
            if (!b.busy && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
        if (best) {
            best->busy = true;
            return best->p;
        }
        void *p = nullptr;
        char err[256] = {0};
        if (dcma_b200_host_alloc(&p, bytes, err, sizeof err) != DCMA_B200_OK) throw std::runtime_error(std::string("dcma_b200: ") + err);
        bufs_.push_back(Buf{p, bytes, true});
        return p;
    }
    void release(void *p) {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto &b : bufs_)
            if (b.p == p) b.busy = false;
    }
};
inline PinnedPool &pool() {
    static PinnedPool p;
    return p;
}
template <class T>
class Pinned {  // RAII lease of a pooled pinned buffer of n elements
    T *p_ = nullptr;
    size_t n_ = 0;

  public:
    Pinned() = default;
    explicit Pinned(size_t n) : p_(static_cast<T *>(pool().acquire(std::max<size_t>(n, 1) * sizeof(T)))), n_(n) {}
    Pinned(const Pinned &) = delete;
    Pinned &operator=(const Pinned &) = delete;
    Pinned(Pinned &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; }
    Pinned &operator=(Pinned &&o) noexcept {
        if (this != &o) {
            if (p_) pool().release(p_);
            p_ = o.p_;
            n_ = o.n_;
            o.p_ = nullptr;
        }
        return *this;
    }
    ~Pinned() {
        if (p_) pool().release(p_);
    }
    T *data() const { return p_; }
    size_t size() const { return n_; }
};

// The regular grid a stack occupies, slice index = LIST order (the order marshal_collection_to_volume uses,
// reference Alignment_Demons.h:139-180).  false: not a regular stack, or not equally spaced in list order.
template <class T>
bool grid_of(const planar_image_collection<T, double> &coll, dcma_b200_grid &g) {
    if (coll.images.empty()) return false;
    auto &nc = const_cast<planar_image_collection<T, double> &>(coll);
    std::list<std::reference_wrapper<planar_image<T, double>>> refs;
    for (auto &img : nc.images) refs.push_back(std::ref(img));
    if (!Images_Form_Regular_Grid(refs)) return false;
    const auto &a = coll.images.front();
    for (const auto &img : coll.images)
        if (img.rows != a.rows || img.columns != a.columns || img.channels != a.channels) return false;
    const auto o = a.position(0, 0);
    auto ez = a.ortho_unit();
    const int64_t n = static_cast<int64_t>(coll.images.size());
    double dz = a.pxl_dz;
    if (n > 1) {
        const auto step = std::next(coll.images.begin())->position(0, 0) - o;
        const double len = step.length();
        if (!(len > 0.0)) return false;
        if (step.Dot(ez) < 0.0) ez = ez * -1.0;  // the list runs against ortho_unit: slices still ascend along ez
        dz = len;
        int64_t i = 0;
        for (const auto &img : coll.images) {
            const auto d = img.position(0, 0) - (o + ez * (dz * static_cast<double>(i)));
            if (d.length() > 1e-6 * std::max(1.0, dz) * 100.0) return false;
            ++i;
        }
    }
    g.geom.slices = n;
    g.geom.rows = a.rows;
    g.geom.cols = a.columns;
    g.geom.channels = a.channels;
    g.geom.pxl_dx = a.pxl_dx;
    g.geom.pxl_dy = a.pxl_dy;
    g.geom.pxl_dz = dz;
    const auto ex = a.row_unit.unit(), ey = a.col_unit.unit();
    g.origin[0] = o.x; g.origin[1] = o.y; g.origin[2] = o.z;
    g.ex[0] = ex.x; g.ex[1] = ex.y; g.ex[2] = ex.z;
    g.ey[0] = ey.x; g.ey[1] = ey.y; g.ey[2] = ey.z;
    g.ez[0] = ez.x; g.ez[1] = ez.y; g.ez[2] = ez.z;
    return true;
}

// image stack -> dense [z][y][x][c] at dst (slice-parallel; memcpy when the image's own layout is that layout)
template <class T>
void marshal_into(const planar_image_collection<T, double> &coll, T *dst) {
    const auto &a = coll.images.front();
    const int64_t R = a.rows, C = a.columns, CH = a.channels;
    std::vector<const planar_image<T, double> *> imgs;
    for (const auto &img : coll.images) {
        if (img.rows != R || img.columns != C || img.channels != CH)
            throw std::invalid_argument("Image collection has inconsistent rows/columns/channels and cannot be marshaled as a rectilinear volume");
        imgs.push_back(&img);
    }
    const bool dense_layout = (a.index(0, 0, 0) == 0) && (CH == 1 || a.index(0, 0, 1) == 1) && (C == 1 || a.index(0, 1, 0) == CH) &&
                              (R == 1 || a.index(1, 0, 0) == C * CH);
    const size_t per = static_cast<size_t>(R * C * CH);
    parallel_for(static_cast<int64_t>(imgs.size()), [&](int64_t z) {
        const auto &img = *imgs[static_cast<size_t>(z)];
        T *d = dst + static_cast<size_t>(z) * per;
        if (dense_layout && img.data.size() == per) {
            std::memcpy(d, img.data.data(), per * sizeof(T));
        } else {
            for (int64_t y = 0; y < R; ++y)
                for (int64_t x = 0; x < C; ++x)
                    for (int64_t c = 0; c < CH; ++c) d[(y * C + x) * CH + c] = img.value(y, x, c);
        }
    });
}

// dense [z][y][x][c] -> one image per slice with the geometry and metadata of the matching slice of `geometry`
// (marshal_volume_to_collection, reference Alignment_Demons.h:182-204), slice-parallel
// Two steps, because the first one does not depend on the voxel values: building the images (allocation and first touch
// of their buffers -- 1.6 GB of fresh pages for a 512x512x256 field, the most expensive host step of a registration) can
// run while the GPU iterates; only the copy has to wait for the result.
template <class T, class G>
planar_image_collection<T, double> collection_like(int64_t slices, int64_t R, int64_t C, int64_t CH,
                                                   const planar_image_collection<G, double> &geometry) {
    if (static_cast<int64_t>(geometry.images.size()) < slices)
        throw std::runtime_error("Requested image index not present, unable to continue");
    planar_image_collection<T, double> out;
    std::vector<planar_image<T, double> *> imgs;
    std::vector<const planar_image<G, double> *> refs;
    auto it = geometry.images.begin();
    for (int64_t z = 0; z < slices; ++z, ++it) {
        out.images.emplace_back();
        imgs.push_back(&out.images.back());
        refs.push_back(&*it);
    }
    parallel_for(slices, [&](int64_t z) {
        auto &img = *imgs[static_cast<size_t>(z)];
        const auto &ref = *refs[static_cast<size_t>(z)];
        img.init_orientation(ref.row_unit, ref.col_unit);
        img.init_buffer(R, C, CH);
        img.init_spatial(ref.pxl_dx, ref.pxl_dy, ref.pxl_dz, ref.anchor, ref.offset);
        img.metadata = ref.metadata;
    });
    return out;
}
template <class T>
void fill_from_dense(planar_image_collection<T, double> &out, const T *src, int64_t R, int64_t C, int64_t CH) {
    std::vector<planar_image<T, double> *> imgs;
    for (auto &img : out.images) imgs.push_back(&img);
    const size_t per = static_cast<size_t>(R * C * CH);
    parallel_for(static_cast<int64_t>(imgs.size()), [&](int64_t z) {
        auto &img = *imgs[static_cast<size_t>(z)];
        const T *s = src + static_cast<size_t>(z) * per;
        const bool dense_layout = (img.index(0, 0, 0) == 0) && (CH == 1 || img.index(0, 0, 1) == 1) &&
                                  (C == 1 || img.index(0, 1, 0) == CH) && (R == 1 || img.index(1, 0, 0) == C * CH);
        if (dense_layout && img.data.size() == per) {
            std::memcpy(img.data.data(), s, per * sizeof(T));
        } else {
            for (int64_t y = 0; y < R; ++y)
                for (int64_t x = 0; x < C; ++x)
                    for (int64_t c = 0; c < CH; ++c) img.reference(y, x, c) = s[(y * C + x) * CH + c];
        }
    });
}
template <class T, class G>
planar_image_collection<T, double> collection_from_dense(const T *src, int64_t slices, int64_t R, int64_t C, int64_t CH,
                                                         const planar_image_collection<G, double> &geometry) {
    auto out = collection_like<T>(slices, R, C, CH, geometry);
    fill_from_dense(out, src, R, C, CH);
    return out;
}

inline void check_rc(int rc, const char *err) {
    if (rc != DCMA_B200_OK) throw std::runtime_error(std::string("dcma_b200: ") + err);
}

}  // namespace detail
}  // namespace dcma_b200_host
