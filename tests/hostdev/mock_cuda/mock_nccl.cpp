// TEST INFRASTRUCTURE: the eight NCCL entry points csrc/ecmgpu.cu resolves through dlopen("libnccl.so.2"), implemented
// for ranks that are THREADS of one process (tests/test_mock_glue.py): a send copies its message into a mailbox keyed by
// (communicator id, source, destination), a receive waits for it.  Built with -Wl,-soname,libnccl.so.2 and loaded into
// the test process before ecmgpu_comm_init, so the product's dlopen finds it by name.  Nothing in the product links it.
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {
struct Comm { int id, rank, n; };
std::mutex g_m;
std::condition_variable g_cv;
std::map<std::tuple<int, int, int>, std::deque<std::vector<char>>> g_box;  // (id, src, dst) -> messages in order
int g_next_id = 1;
struct Op { bool send; void* buf; size_t n; int peer; Comm* c; };
thread_local bool t_group = false;
thread_local std::vector<Op> t_ops;

void run(const Op& o) {
    if (o.send) {
        std::lock_guard<std::mutex> lk(g_m);
        g_box[{o.c->id, o.c->rank, o.peer}].emplace_back((char*)o.buf, (char*)o.buf + o.n);
        g_cv.notify_all();
    } else {
        std::unique_lock<std::mutex> lk(g_m);
        auto& q = g_box[{o.c->id, o.peer, o.c->rank}];
        g_cv.wait(lk, [&] { return !q.empty(); });
        memcpy(o.buf, q.front().data(), std::min(o.n, q.front().size()));
        q.pop_front();
    }
}
int post(const Op& o) {
    if (t_group) t_ops.push_back(o);
    else run(o);
    return 0;
}
}  // namespace

extern "C" {
int ncclGetUniqueId(void* id) {
    std::lock_guard<std::mutex> lk(g_m);
    memset(id, 0, 128);
    const int v = g_next_id++;
    memcpy(id, &v, sizeof(v));
    return 0;
}
struct UniqueId { char internal[128]; };
int ncclCommInitRank(void** comm, int n, UniqueId id, int rank) {
    int v;
    memcpy(&v, id.internal, sizeof(v));
    *comm = new Comm{v, rank, n};
    return 0;
}
int ncclCommDestroy(void* comm) { delete (Comm*)comm; return 0; }
int ncclSend(const void* buf, size_t count, int /*ncclInt8*/, int peer, void* comm, void* /*stream*/) { return post(Op{true, (void*)buf, count, peer, (Comm*)comm}); }
int ncclRecv(void* buf, size_t count, int, int peer, void* comm, void*) { return post(Op{false, buf, count, peer, (Comm*)comm}); }
int ncclGroupStart() { t_group = true; return 0; }
int ncclGroupEnd() {  // every send of the group first, then the receives: no order of calls can deadlock
    t_group = false;
    for (const Op& o : t_ops) if (o.send) run(o);
    for (const Op& o : t_ops) if (!o.send) run(o);
    t_ops.clear();
    return 0;
}
const char* ncclGetErrorString(int) { return "mock NCCL error"; }
}
