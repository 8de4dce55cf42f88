#include "gpu_bridge.h"

#include <array>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace ipxb200 {

namespace {

struct ModelEntry {
    ipxgpu_ctx* ctx = nullptr;
    unsigned long long generation = 0;
    unsigned long long last_use = 0;
    const double* values = nullptr;
    ipx::Int entries = 0, rows = 0, cols = 0;
    unsigned long long fingerprint = 0;
    unsigned long long light_fingerprint = 0;  // 256 samples: checked on every MultiplyAdd
    const double* hint_W = nullptr;
    bool hint_weights = false, hint_diag = false;
};

struct Tables {
    std::mutex mutex;
    std::unordered_map<const ipx::Model*, ModelEntry> models;
    std::unordered_map<const ipx::Model*, ModelEntry> groups;  // multi-GPU groups (IPXGPU_NGPUS)
    std::unordered_map<const ipx::LinearOperator*, OperatorRecord> operators;
    std::unordered_map<const ipxgpu_ctx*, std::array<const void*, (size_t)StateSlot::kCount>> owners;
    unsigned long long generation = 0, clock = 0;
    ~Tables() {
        for (auto& kv : models)
            if (kv.second.ctx) ipxgpu_destroy(kv.second.ctx);
        for (auto& kv : groups)
            if (kv.second.ctx) ipxgpu_destroy(kv.second.ctx);
    }
};

Tables& tables() {
    static Tables t;
    return t;
}

// Content fingerprint over a strided sample of AI (<= 64k entries): guards
// against a new Model whose arrays were allocated at a recycled address.
unsigned long long Fingerprint(const ipx::SparseMatrix& AI, ipx::Int samples = 65536) {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) {
        h ^= v;
        h *= 1099511628211ull;
    };
    const ipx::Int nz = AI.entries(), nc = AI.cols();
    const ipx::Int step = nz > samples ? nz / samples : 1;
    for (ipx::Int p = 0; p < nz; p += step) {
        unsigned long long bits;
        const double v = AI.value(p);
        static_assert(sizeof bits == sizeof v, "double must be 64-bit");
        std::memcpy(&bits, &v, sizeof bits);
        mix(bits);
        mix((unsigned long long)AI.index(p));
    }
    const ipx::Int cstep = nc > samples ? nc / samples : 1;
    for (ipx::Int j = 0; j <= nc; j += cstep) mix((unsigned long long)AI.colptr()[j]);
    return h;
}

ModelEntry* EntryOfContext(ipxgpu_ctx* ctx) {
    for (auto& kv : tables().models)
        if (kv.second.ctx == ctx) return &kv.second;
    return nullptr;
}

// The CUDA context is created on a helper thread from the moment the drop-in build of IPX is
// loaded, so that it is ready (or well under way) when the first KKTSolverDiag::Factorize needs
// it. IPXGPU_EAGER_INIT=0 defers it to the first context.
const bool g_eager_init = [] {
    const char* env = std::getenv("IPXGPU_EAGER_INIT");
    if (env && std::atoi(env) == 0) return false;
    const char* dev = std::getenv("IPXGPU_DEVICE");
    ipxgpu_warmup(dev ? std::atoi(dev) : 0);
    return true;
}();

size_t MaxContexts() {
    const char* env = std::getenv("IPXGPU_MAX_CONTEXTS");
    const long v = env ? std::atol(env) : 4;
    return v < 1 ? 1 : (size_t)v;
}

}  // namespace

void Check(int rc) {
    if (rc == IPXGPU_OK) return;
    if (rc == IPXGPU_ERR_OUT_OF_MEMORY) throw std::bad_alloc();
    throw std::runtime_error(std::string("ipxgpu: ") + ipxgpu_last_error());
}

ContextRef ContextFor(const ipx::Model& model) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    const ipx::SparseMatrix& AI = model.AI();
    const unsigned long long fp = Fingerprint(AI);
    auto it = t.models.find(&model);
    if (it != t.models.end()) {
        ModelEntry& e = it->second;
        if (e.ctx && e.values == AI.values() && e.entries == AI.entries() &&
            e.rows == model.rows() && e.cols == model.cols() && e.fingerprint == fp) {
            e.last_use = ++t.clock;
            return ContextRef{e.ctx, e.generation};
        }
        if (e.ctx) {
            t.owners.erase(e.ctx);
            ipxgpu_destroy(e.ctx);
        }
        t.models.erase(it);
    }
    // Model has no destructor hook: evict the least recently used context
    // once the cache is full. Operators notice through the generation.
    while (t.models.size() >= MaxContexts()) {
        auto victim = t.models.begin();
        for (auto jt = t.models.begin(); jt != t.models.end(); ++jt)
            if (jt->second.last_use < victim->second.last_use) victim = jt;
        if (victim->second.ctx) {
            t.owners.erase(victim->second.ctx);
            ipxgpu_destroy(victim->second.ctx);
        }
        t.models.erase(victim);
    }
    ipxgpu_options opt;
    ipxgpu_default_options(&opt);
    ipxgpu_ctx* ctx = nullptr;
    Check(ipxgpu_create(&ctx, model.rows(), model.cols(), AI.colptr(), AI.rowidx(), AI.values(),
                        &opt));
    ModelEntry& e = t.models[&model];
    e.ctx = ctx;
    e.generation = ++t.generation;
    e.last_use = ++t.clock;
    e.values = AI.values();
    e.entries = AI.entries();
    e.rows = model.rows();
    e.cols = model.cols();
    e.fingerprint = fp;
    e.light_fingerprint = Fingerprint(AI, 256);
    return ContextRef{e.ctx, e.generation};
}

ContextRef CurrentContext(const ipx::Model& model) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    auto it = t.models.find(&model);
    if (it == t.models.end() || !it->second.ctx) return ContextRef{};
    it->second.last_use = ++t.clock;
    return ContextRef{it->second.ctx, it->second.generation};
}

ipxgpu_ctx* ContextOfMatrix(const ipx::SparseMatrix& A) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    for (auto& kv : t.models) {
        ModelEntry& e = kv.second;
        // AI has rows + cols columns (structural + slack)
        if (!e.ctx || e.values != A.values() || e.entries != A.entries() || e.rows != A.rows() ||
            e.rows + e.cols != A.cols())
            continue;
        if (e.light_fingerprint != Fingerprint(A, 256)) return nullptr;  // recycled address, other content
        e.last_use = ++t.clock;
        return e.ctx;
    }
    return nullptr;
}

int GroupSize() {
    static const int g = [] {
        const char* env = std::getenv("IPXGPU_NGPUS");
        const long v = env ? std::atol(env) : 1;
        return (int)(v < 1 ? 1 : (v > 16 ? 16 : v));
    }();
    return g;
}

ContextRef GroupContextFor(const ipx::Model& model) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    const ipx::SparseMatrix& AI = model.AI();
    const unsigned long long fp = Fingerprint(AI);
    auto it = t.groups.find(&model);
    if (it != t.groups.end()) {
        ModelEntry& e = it->second;
        if (e.ctx && e.values == AI.values() && e.entries == AI.entries() &&
            e.rows == model.rows() && e.cols == model.cols() && e.fingerprint == fp) {
            e.last_use = ++t.clock;
            return ContextRef{e.ctx, e.generation};
        }
        if (e.ctx) ipxgpu_destroy(e.ctx);
        t.groups.erase(it);
    }
    // one group at a time: a group holds the matrix on every GPU
    for (auto& kv : t.groups)
        if (kv.second.ctx) ipxgpu_destroy(kv.second.ctx);
    t.groups.clear();
    ipxgpu_options opt;
    ipxgpu_default_options(&opt);
    ipxgpu_ctx* ctx = nullptr;
    Check(ipxgpu_create_group(&ctx, model.rows(), model.cols(), AI.colptr(), AI.rowidx(),
                              AI.values(), &opt, GroupSize(), nullptr));
    ModelEntry& e = t.groups[&model];
    e.ctx = ctx;
    e.generation = ++t.generation;
    e.last_use = ++t.clock;
    e.values = AI.values();
    e.entries = AI.entries();
    e.rows = model.rows();
    e.cols = model.cols();
    e.fingerprint = fp;
    e.light_fingerprint = Fingerprint(AI, 256);
    return ContextRef{e.ctx, e.generation};
}

ContextRef CurrentGroupContext(const ipx::Model& model) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    auto it = t.groups.find(&model);
    if (it == t.groups.end() || !it->second.ctx) return ContextRef{};
    return ContextRef{it->second.ctx, it->second.generation};
}

OperatorRecord& RecordOf(const ipx::LinearOperator* op) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    return t.operators[op];
}

OperatorRecord* FindRecord(const ipx::LinearOperator* op) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    auto it = t.operators.find(op);
    return it == t.operators.end() ? nullptr : &it->second;
}

void Forget(const ipx::LinearOperator* op) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    t.operators.erase(op);
}

bool StillCurrent(const OperatorRecord& rec) {
    if (!rec.model || !rec.ref.ctx) return false;
    const ContextRef cur = CurrentContext(*rec.model);
    return cur.ctx == rec.ref.ctx && cur.generation == rec.ref.generation;
}

void ClaimState(ipxgpu_ctx* ctx, StateSlot slot, const void* owner) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    auto it = t.owners.find(ctx);
    if (it == t.owners.end()) it = t.owners.emplace(ctx, decltype(t.owners)::mapped_type{}).first;
    it->second[(size_t)slot] = owner;
}

bool OwnsState(ipxgpu_ctx* ctx, StateSlot slot, const void* owner) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    auto it = t.owners.find(ctx);
    return it != t.owners.end() && it->second[(size_t)slot] == owner;
}

void EnsurePrimed(OperatorRecord& rec, const ipx::LinearOperator* op) {
    StateSlot slot;
    switch (rec.kind) {
        case OperatorKind::kNormal: slot = StateSlot::kWeights; break;
        case OperatorKind::kDiagonal: slot = StateSlot::kDiagonal; break;
        case OperatorKind::kSplit: slot = StateSlot::kSplit; break;
        default: return;
    }
    if (OwnsState(rec.ref.ctx, slot, op)) return;
    if (!rec.reprime)
        throw std::logic_error("another operator instance on this model has replaced the device "
                               "state of this one; prepare it again");
    rec.reprime();
}

void SetResidentHint(ipxgpu_ctx* ctx, const double* W) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    if (ModelEntry* e = EntryOfContext(ctx)) {
        e->hint_W = W;
        e->hint_weights = true;
        e->hint_diag = true;
    }
}

bool ConsumeWeightsHint(ipxgpu_ctx* ctx, const double* W) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    ModelEntry* e = EntryOfContext(ctx);
    if (!e || !e->hint_weights || e->hint_W != W) return false;
    e->hint_weights = false;
    return true;
}

bool ConsumeDiagonalHint(ipxgpu_ctx* ctx, const double* W) {
    Tables& t = tables();
    std::lock_guard<std::mutex> lock(t.mutex);
    ModelEntry* e = EntryOfContext(ctx);
    if (!e || !e->hint_diag || e->hint_W != W) return false;
    e->hint_diag = false;
    return true;
}

}  // namespace ipxb200
