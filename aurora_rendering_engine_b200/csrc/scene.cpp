// scene.cpp — scene compiler: validation, plane-form records, parallelogram fusion, brute list, binned-SAH BVH2.
#include "scene.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <type_traits>
#include <utility>
#include <unordered_map>

#include "philox.cuh"

namespace areb {

namespace {

struct D3 {
	double x, y, z;
};
inline D3 d3(const double *p) { return { p[0], p[1], p[2] }; }
inline D3 operator+(D3 a, D3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline D3 operator-(D3 a, D3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline D3 operator*(double s, D3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 cross(D3 a, D3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double len(D3 a) { return std::sqrt(dot(a, a)); }
inline bool near_zero(D3 a) {  // are::Vec3::near_zero, src/basic/vec3.cpp:143-146
	const double s = 1e-8;
	return std::fabs(a.x) < s && std::fabs(a.y) < s && std::fabs(a.z) < s;
}

// Host threads for the data-parallel phases of the scene compiler (set_build_threads overrides the hardware count).
std::atomic<int> g_build_threads{ 0 };
inline unsigned hw_threads() {
	const int forced = g_build_threads.load(std::memory_order_relaxed);
	return forced > 0 ? (unsigned)forced : std::thread::hardware_concurrency();
}
inline unsigned host_threads() { return std::max(1u, std::min(hw_threads(), 64u)); }
// fn(begin, end) over [0,n) in contiguous chunks, one per thread; results must not depend on the chunking
template <typename F>
void parallel_chunks(size_t n, size_t min_chunk, F fn) {
	const unsigned T = (unsigned)std::min<size_t>(host_threads(), std::max<size_t>(1, n / std::max<size_t>(1, min_chunk)));
	if (T <= 1) { fn((size_t)0, n); return; }
	std::vector<std::thread> th;
	for (unsigned t = 0; t < T; ++t) th.emplace_back([=] { fn(n * t / T, n * (t + 1) / T); });
	for (auto &x : th) x.join();
}

struct Box {
	double lo[3], hi[3];
	void reset() {
		for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<double>::infinity(); hi[k] = -lo[k]; }
	}
	void grow(D3 p) {
		const double v[3] = { p.x, p.y, p.z };
		for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
	}
	void grow(const Box &b) {
		for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
	}
	double area() const {
		double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
		if (dx < 0 || dy < 0 || dz < 0) return 0.0;
		return 2.0 * (dx * dy + dy * dz + dz * dx);
	}
};

// plane form of the parallelogram / triangle (Q,u,v)
HotPrim plane_form(D3 Q, D3 u, D3 v) {
	D3 N = cross(u, v);
	double nn = dot(N, N);
	D3 n = (1.0 / std::sqrt(nn)) * N;
	D3 A = (1.0 / nn) * cross(v, N), B = (1.0 / nn) * cross(N, u);
	HotPrim h;
	h.r0 = { (float)n.x, (float)n.y, (float)n.z, (float)dot(n, Q) };
	h.r1 = { (float)A.x, (float)A.y, (float)A.z, (float)dot(A, Q) };
	h.r2 = { (float)B.x, (float)B.y, (float)B.z, (float)dot(B, Q) };
	return h;
}
HotPrim sphere_form(D3 c, double r) {
	HotPrim h;
	h.r0 = { (float)c.x, (float)c.y, (float)c.z, (float)r };
	h.r1 = { (float)(r * r), 0.f, 0.f, 0.f };
	h.r2 = { 0.f, 0.f, 0.f, 0.f };
	return h;
}

enum HotKind { HK_BOX = 0, HK_QUAD = 1, HK_TRI = 2, HK_SPHERE = 3, HK_KINDS = 4 };  // emission order inside a range
struct HotItem {
	int kind;
	HotPrim rec, rec2;  // rec2: second slot of a box
	HotIds ids;
	Box box;
	double c[3];  // centroid
	D3 gQ, gu, gv;  // parallelogram geometry (quads / fused pairs), for box detection
	bool dead;  // absorbed into a box (no default member initialiser: the struct stays trivial, see HotVec)
};
// std::vector whose resize() default-initialises, i.e. leaves trivial elements untouched: the 260-byte hot items of a
// million-primitive scene are then first touched by the threads that fill them, not zero-filled by one thread first.
template <typename T>
struct DefaultInitAlloc : std::allocator<T> {
	template <typename U> struct rebind { using other = DefaultInitAlloc<U>; };
	template <typename U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
	template <typename U, typename... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
typedef std::vector<HotItem, DefaultInitAlloc<HotItem>> HotVec;
inline void emit_item(const HotItem &it, std::vector<HotPrim> &prims, std::vector<HotIds> &ids) {
	prims.push_back(it.rec);
	ids.push_back(it.ids);
	if (it.kind == HK_BOX) { prims.push_back(it.rec2); ids.push_back(it.ids); }
}

struct VKey {
	uint64_t a[3];
	bool operator==(const VKey &o) const { return a[0] == o.a[0] && a[1] == o.a[1] && a[2] == o.a[2]; }
};
inline VKey vkey(D3 p) {
	VKey k;
	double v[3] = { p.x == 0.0 ? 0.0 : p.x, p.y == 0.0 ? 0.0 : p.y, p.z == 0.0 ? 0.0 : p.z };  // -0 -> +0
	std::memcpy(k.a, v, sizeof v);
	return k;
}
struct EKey {
	VKey p, q;
	bool operator==(const EKey &o) const { return p == o.p && q == o.q; }
};
struct EHash {
	size_t operator()(const EKey &e) const {
		uint64_t h = 0x9E3779B97F4A7C15ull;
		const uint64_t *w = e.p.a;
		for (int i = 0; i < 3; ++i) h = (h ^ w[i]) * 0xD6E8FEB86659FD93ull;
		w = e.q.a;
		for (int i = 0; i < 3; ++i) h = (h ^ w[i]) * 0xD6E8FEB86659FD93ull;
		return (size_t)(h ^ (h >> 32));
	}
};
inline bool vless(const VKey &x, const VKey &y) { return std::lexicographical_compare(x.a, x.a + 3, y.a, y.a + 3); }
inline EKey ekey(D3 p, D3 q) {
	VKey a = vkey(p), b = vkey(q);
	if (vless(b, a)) std::swap(a, b);
	return { a, b };
}

// ---- BVH builder ------------------------------------------------------------------------------------
struct ChildRef {
	int ref, meta;
	Box box;
	int depth;  // height of the subtree
};

// Binned-SAH BVH2 with exactly one hot item per leaf.  A subtree over k items therefore has k-1 inner nodes and its
// items fill a contiguous slot range, so every subtree's place in the depth-first node / primitive arrays is known
// before it is built: subtrees are built by independent threads straight into the preallocated output (no stitching),
// and the result is identical to the serial build whatever the thread count.
struct Builder {
	// what the builder touches per item, 80 bytes, physically partitioned with the ranges so that every subtree scans
	// contiguous memory (the full HotItem is 256 bytes and would be visited through a permutation)
	struct Ref {
		Box box;
		double c[3];  // centroid
		int item;
	};
	const HotVec &items;
	std::vector<Ref> refs;
	CompiledScene &out;
	bool any_box = false;
	int spawn_depth = 0;  // subtrees above this depth (and big enough) hand their left half to a new thread
	Builder(const HotVec &it, CompiledScene &o) : items(it), out(o) {
		refs.resize(items.size());
		size_t slots = 0;
		for (size_t i = 0; i < refs.size(); ++i) {
			refs[i].box = items[i].box;
			for (int k = 0; k < 3; ++k) refs[i].c[k] = items[i].c[k];
			refs[i].item = (int)i;
			any_box |= items[i].kind == HK_BOX;
			slots += items[i].kind == HK_BOX ? 2 : 1;
		}
		out.nodes.assign(items.empty() ? 0 : items.size() - 1, BvhNode());
		out.bvh_prims.assign(slots, HotPrim());
		out.bvh_ids.assign(slots, HotIds{ -1, -1 });
		const unsigned hw = std::max(1u, hw_threads());
		while ((1u << spawn_depth) < hw && spawn_depth < 8) ++spawn_depth;
		if (hw > 1) spawn_depth += 2;  // 4 tasks per core: SAH splits are uneven
		else spawn_depth = 0;
	}
	// Child boxes are stored as (centre, half-extent) per axis: the slab distances are then centre*inv -/+ half*|inv|,
	// three FMA-pipe operations per axis and no per-axis min/max (the ALU pipe is what binds the traversal kernels).
	static void put_box(f4 &bxy, float &zc, float &zh, const Box &b) {
		float c[3], h[3];
		for (int k = 0; k < 3; ++k) {
			// pad outwards: fp32 rounding of the bounds and of the slab arithmetic must never cut a primitive off
			const double m = std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k]));
			const double pad = 1e-5 * (m + (b.hi[k] - b.lo[k])) + 1e-7;
			const double lo = b.lo[k] - pad, hi = b.hi[k] + pad;
			c[k] = (float)(0.5 * (lo + hi));
			h[k] = std::nextafter((float)std::max(hi - (double)c[k], (double)c[k] - lo), std::numeric_limits<float>::infinity());
		}
		bxy = { c[0], h[0], c[1], h[1] };
		zc = c[2];
		zh = h[2];
	}
	int slots_of(int lo, int hi) const {
		if (!any_box) return hi - lo;
		int n = 0;
		for (int i = lo; i < hi; ++i) n += items[refs[i].item].kind == HK_BOX ? 2 : 1;
		return n;
	}
	ChildRef make_leaf(int at, int slot) {  // ref = ~(slot | kind << 29)
		ChildRef c;
		const HotItem &it = items[refs[at].item];
		out.bvh_prims[slot] = it.rec;
		out.bvh_ids[slot] = it.ids;
		if (it.kind == HK_BOX) { out.bvh_prims[slot + 1] = it.rec2; out.bvh_ids[slot + 1] = it.ids; }
		c.box = it.box;
		c.ref = ~(slot | (it.kind << 29));
		c.meta = 0;
		c.depth = 0;
		return c;
	}
	// Subtree over refs[lo,hi): inner nodes go to out.nodes[node_base ...), items to slots [slot_base ...).
	ChildRef build(int lo, int hi, int depth, int node_base, int slot_base) {
		const int n = hi - lo;
		if (n <= 1) return make_leaf(lo, slot_base);
		Box cb;
		cb.reset();
		for (int i = lo; i < hi; ++i) cb.grow(D3{ refs[i].c[0], refs[i].c[1], refs[i].c[2] });
		int mid = -1;
		// SAH splits can be arbitrarily uneven; median splits halve.  SAH is used only while a median-split subtree under the
		// worse child (n - 1 items, height ceil(log2(n - 1))) would still fit the traversal stack, so the finished tree never
		// exceeds ARE_BVH_STACK levels whatever the scene (skewed scenes used to fail the commit instead).
		int lg = 0;
		while ((1 << lg) < n) ++lg;
		if (depth + lg + 2 <= ARE_BVH_STACK) {
			const int NB = 16;
			double best_cost = std::numeric_limits<double>::infinity();
			int best_axis = -1, best_split = -1;
			// one pass fills the bins of all three axes
			Box bb[3][NB];
			int bc[3][NB];
			double scale[3];
			bool live[3];
			for (int ax = 0; ax < 3; ++ax) {
				const double ext = cb.hi[ax] - cb.lo[ax];
				live[ax] = ext > 0.0;
				scale[ax] = live[ax] ? NB / ext : 0.0;
				for (int b = 0; b < NB; ++b) { bb[ax][b].reset(); bc[ax][b] = 0; }
			}
			for (int i = lo; i < hi; ++i) {
				const Ref &r = refs[i];
				for (int ax = 0; ax < 3; ++ax) {
					if (!live[ax]) continue;
					const int b = std::min(NB - 1, (int)((r.c[ax] - cb.lo[ax]) * scale[ax]));
					bb[ax][b].grow(r.box);
					bc[ax][b]++;
				}
			}
			for (int ax = 0; ax < 3; ++ax) {
				if (!live[ax]) continue;
				double right_area[NB];
				int right_cnt[NB];
				Box acc;
				acc.reset();
				int c = 0;
				for (int b = NB - 1; b > 0; --b) {
					acc.grow(bb[ax][b]);
					c += bc[ax][b];
					right_area[b] = acc.area();
					right_cnt[b] = c;
				}
				acc.reset();
				c = 0;
				for (int b = 0; b < NB - 1; ++b) {
					acc.grow(bb[ax][b]);
					c += bc[ax][b];
					if (c == 0 || right_cnt[b + 1] == 0) continue;
					double cost = acc.area() * c + right_area[b + 1] * right_cnt[b + 1];
					if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; }
				}
			}
			if (best_axis >= 0) {
				const double sc = scale[best_axis], clo = cb.lo[best_axis];
				auto it = std::partition(refs.begin() + lo, refs.begin() + hi, [&](const Ref &r) {
					int b = std::min(NB - 1, (int)((r.c[best_axis] - clo) * sc));
					return b <= best_split;
				});
				mid = (int)(it - refs.begin());
				if (mid == lo || mid == hi) mid = -1;
			}
		}
		if (mid < 0) {  // coincident centroids or depth guard: median split on the widest axis
			int ax = 0;
			for (int k = 1; k < 3; ++k)
				if (cb.hi[k] - cb.lo[k] > cb.hi[ax] - cb.lo[ax]) ax = k;
			mid = lo + n / 2;
			std::nth_element(refs.begin() + lo, refs.begin() + mid, refs.begin() + hi,
				[&](const Ref &a, const Ref &b) { return a.c[ax] != b.c[ax] ? a.c[ax] < b.c[ax] : a.item < b.item; });
		}
		// depth-first layout: this node, the left subtree's (mid-lo)-1 nodes, then the right subtree's
		const int me = node_base, left_nodes = node_base + 1, right_nodes = node_base + (mid - lo);
		const int right_slots = slot_base + slots_of(lo, mid);
		ChildRef l, r;
		if (depth < spawn_depth && n >= 16384) {
			std::thread left([&] { l = build(lo, mid, depth + 1, left_nodes, slot_base); });
			r = build(mid, hi, depth + 1, right_nodes, right_slots);
			left.join();
		} else {
			l = build(lo, mid, depth + 1, left_nodes, slot_base);
			r = build(mid, hi, depth + 1, right_nodes, right_slots);
		}
		BvhNode nd;
		put_box(nd.b0, nd.b2.x, nd.b2.y, l.box);
		put_box(nd.b1, nd.b2.z, nd.b2.w, r.box);
		nd.child[0] = l.ref; nd.child[1] = r.ref;
		nd.meta[0] = l.meta; nd.meta[1] = r.meta;
		out.nodes[me] = nd;
		ChildRef c;
		c.ref = me;
		c.meta = 0;
		c.box = l.box;
		c.box.grow(r.box);
		c.depth = 1 + std::max(l.depth, r.depth);
		return c;
	}
};

// ---- BVH2 -> uncompressed 4-wide collapse (Bvh4Node, dev_types.h) -----------------------------------------
// A 4-wide node starts from a BVH2 node's two children and absorbs the inner child of largest surface area until it
// has four (or only leaves are left).  Child boxes are the (centre, half-extent) floats the BVH2 parent already holds,
// copied bit for bit, so a BVH4 traversal accepts exactly the boxes the BVH2 traversal accepts.
struct Bvh4Builder {
	const CompiledScene &src;
	std::vector<Bvh4Node> &out;
	int depth = 0;
	struct Kid {
		int ref;
		float c[3], h[3];
		float area() const { return h[0] * h[1] + h[1] * h[2] + h[2] * h[0]; }
	};
	static void kids_of(const BvhNode &n, Kid &a, Kid &b) {
		a.ref = n.child[0]; b.ref = n.child[1];
		a.c[0] = n.b0.x; a.h[0] = n.b0.y; a.c[1] = n.b0.z; a.h[1] = n.b0.w; a.c[2] = n.b2.x; a.h[2] = n.b2.y;
		b.c[0] = n.b1.x; b.h[0] = n.b1.y; b.c[1] = n.b1.z; b.h[1] = n.b1.w; b.c[2] = n.b2.z; b.h[2] = n.b2.w;
	}
	int build(int node2, int level) {
		depth = std::max(depth, level + 1);
		Kid k[4];
		int n = 2;
		kids_of(src.nodes[(size_t)node2], k[0], k[1]);
		while (n < 4) {
			int best = -1;
			for (int i = 0; i < n; ++i)
				if (k[i].ref >= 0 && (best < 0 || k[i].area() > k[best].area())) best = i;
			if (best < 0) break;
			Kid a, b;
			kids_of(src.nodes[(size_t)k[best].ref], a, b);
			k[best] = a;
			k[n++] = b;
		}
		const int me = (int)out.size();
		out.push_back(Bvh4Node());
		Bvh4Node nd;
		float *cx = &nd.cx.x, *hx = &nd.hx.x, *cy = &nd.cy.x, *hy = &nd.hy.x, *cz = &nd.cz.x, *hz = &nd.hz.x;
		for (int i = 0; i < 4; ++i) {
			if (i < n) {
				cx[i] = k[i].c[0]; hx[i] = k[i].h[0]; cy[i] = k[i].c[1]; hy[i] = k[i].h[1]; cz[i] = k[i].c[2]; hz[i] = k[i].h[2];
				nd.child[i] = k[i].ref >= 0 ? build(k[i].ref, level + 1) : k[i].ref;
			} else {
				cx[i] = cy[i] = cz[i] = 0.f; hx[i] = hy[i] = hz[i] = -1e30f;  // empty slot: near > far for every ray
				nd.child[i] = (int)0x80000000;
			}
			nd.pad_[i] = 0;
		}
		out[(size_t)me] = nd;
		return me;
	}
};

// ---- BVH2 -> compressed 8-wide collapse (WideNode, dev_types.h) -----------------------------------------
// Greedy: a wide node starts from a BVH2 node's two children and keeps opening the inner child with the largest
// surface area until it has eight (or only leaves are left).  Children go to the slot whose octant direction
// (+/-,+/-,+/-) they lie furthest along, so that slot XOR ray-octant orders them front to back during traversal.
struct WideBuilder {
	const CompiledScene &src;   // the BVH2 (nodes + leaf-ordered prims)
	CompiledScene &out;
	int max_depth = 0;
	struct Entry { int ref; float lo[3], hi[3]; };
	WideBuilder(const CompiledScene &s, CompiledScene &o) : src(s), out(o) {}
	static Entry child_of(const BvhNode &n, int which) {
		Entry e;
		e.ref = n.child[which];
		const f4 &bxy = which ? n.b1 : n.b0;
		const float c[3] = { bxy.x, bxy.z, which ? n.b2.z : n.b2.x }, h[3] = { bxy.y, bxy.w, which ? n.b2.w : n.b2.y };
		for (int k = 0; k < 3; ++k) {  // (centre, half-extent) -> conservative bounds
			e.lo[k] = std::nextafter(c[k] - h[k], -std::numeric_limits<float>::infinity());
			e.hi[k] = std::nextafter(c[k] + h[k], std::numeric_limits<float>::infinity());
		}
		return e;
	}
	static double area(const Entry &e) {
		const double dx = (double)e.hi[0] - e.lo[0], dy = (double)e.hi[1] - e.lo[1], dz = (double)e.hi[2] - e.lo[2];
		return dx * dy + dy * dz + dz * dx;
	}
	void build() {
		out.wnodes.clear(); out.wide_prims.clear(); out.wide_ids.clear(); out.wide_kinds.clear();
		if (src.nodes.empty()) return;
		out.wnodes.reserve(src.nodes.size() / 3 + 8);
		out.wide_prims.reserve(src.bvh_prims.size());
		out.wide_ids.reserve(src.bvh_ids.size());
		out.wide_kinds.reserve(src.bvh_prims.size());
		out.wnodes.push_back(WideNode());
		// explicit stack (depth-first): (BVH2 node, wide node index, depth)
		struct Job { int bvh2, wide, depth; };
		std::vector<Job> jobs;
		jobs.push_back({ 0, 0, 1 });
		while (!jobs.empty()) {
			const Job job = jobs.back();
			jobs.pop_back();
			max_depth = std::max(max_depth, job.depth);
			Entry kids[8];
			int nk = 2;
			kids[0] = child_of(src.nodes[job.bvh2], 0);
			kids[1] = child_of(src.nodes[job.bvh2], 1);
			while (nk < 8) {
				int best = -1;
				double best_area = -1.0;
				for (int i = 0; i < nk; ++i)
					if (kids[i].ref >= 0 && area(kids[i]) > best_area) { best_area = area(kids[i]); best = i; }
				if (best < 0) break;
				const BvhNode &open = src.nodes[kids[best].ref];
				kids[best] = child_of(open, 0);
				kids[nk++] = child_of(open, 1);
			}
			// node bounds, quantisation grid
			float lo[3], hi[3];
			for (int k = 0; k < 3; ++k) { lo[k] = kids[0].lo[k]; hi[k] = kids[0].hi[k]; }
			for (int i = 1; i < nk; ++i)
				for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], kids[i].lo[k]); hi[k] = std::max(hi[k], kids[i].hi[k]); }
			int ebits[3];
			double step[3];
			for (int k = 0; k < 3; ++k) {
				const double ext = (double)hi[k] - (double)lo[k];
				int e = -126;
				if (ext > 0.0) {
					int fe;
					std::frexp(ext / 255.0, &fe);  // ext/255 = m * 2^fe, m in [0.5,1)  =>  2^fe >= ext/255
					e = std::max(-126, std::min(126, fe));
				}
				for (;; ++e) {  // the ceil() of the upper bounds must still fit 8 bits
					step[k] = std::ldexp(1.0, e);
					bool fits = true;
					for (int i = 0; i < nk && fits; ++i) fits = std::ceil(((double)kids[i].hi[k] - (double)lo[k]) / step[k]) <= 255.0;
					if (fits || e >= 126) break;
				}
				ebits[k] = e + 127;
			}
			// slot assignment: greedy on cost(child, slot) = (centroid - node centre) . (+/-1, +/-1, +/-1)
			int slot_of[8], child_in[8];
			for (int i = 0; i < 8; ++i) { slot_of[i] = -1; child_in[i] = -1; }
			double cost[8][8];
			for (int i = 0; i < nk; ++i)
				for (int sl = 0; sl < 8; ++sl) {
					double c = 0.0;
					for (int k = 0; k < 3; ++k) {
						const double rel = 0.5 * ((double)kids[i].lo[k] + kids[i].hi[k]) - 0.5 * ((double)lo[k] + hi[k]);
						c += ((sl >> k) & 1) ? rel : -rel;
					}
					cost[i][sl] = c;
				}
			for (int round = 0; round < nk; ++round) {
				int bi = -1, bs = -1;
				double bc = -std::numeric_limits<double>::infinity();
				for (int i = 0; i < nk; ++i) {
					if (slot_of[i] >= 0) continue;
					for (int sl = 0; sl < 8; ++sl)
						if (child_in[sl] < 0 && cost[i][sl] > bc) { bc = cost[i][sl]; bi = i; bs = sl; }
				}
				slot_of[bi] = bs;
				child_in[bs] = bi;
			}
			// emit
			unsigned char meta[8] = { 0 }, q[6][8];
			std::memset(q, 0, sizeof q);
			unsigned imask = 0;
			const int child_base = (int)out.wnodes.size(), prim_base = (int)out.wide_prims.size();
			int n_inner = 0, prim_off = 0;
			for (int sl = 0; sl < 8; ++sl) {
				const int i = child_in[sl];
				if (i < 0) continue;
				const Entry &e = kids[i];
				for (int k = 0; k < 3; ++k) {
					q[k][sl] = (unsigned char)std::max(0.0, std::min(255.0, std::floor(((double)e.lo[k] - (double)lo[k]) / step[k])));
					q[3 + k][sl] = (unsigned char)std::max(0.0, std::min(255.0, std::ceil(((double)e.hi[k] - (double)lo[k]) / step[k])));
				}
				if (e.ref >= 0) {
					imask |= 1u << sl;
					meta[sl] = (unsigned char)(0x20 | (24 + sl));
					++n_inner;
				} else {
					const unsigned code = (unsigned)~e.ref;
					const int slot = (int)(code & 0x1fffffffu), kind = (int)(code >> 29);
					meta[sl] = (unsigned char)(0x20 | prim_off);
					const int span = kind == HK_BOX ? 2 : 1;
					for (int m = 0; m < span; ++m) {
						out.wide_prims.push_back(src.bvh_prims[slot + m]);
						out.wide_ids.push_back(src.bvh_ids[slot + m]);
						out.wide_kinds.push_back((unsigned char)kind);
					}
					prim_off += span;
				}
			}
			out.wnodes.resize(out.wnodes.size() + n_inner);
			int rel = 0;
			for (int sl = 0; sl < 8; ++sl) {
				const int i = child_in[sl];
				if (i < 0 || kids[i].ref < 0) continue;
				jobs.push_back({ kids[i].ref, child_base + rel, job.depth + 1 });
				++rel;
			}
			WideNode w;
			unsigned ew = (unsigned)ebits[0] | (unsigned)ebits[1] << 8 | (unsigned)ebits[2] << 16 | imask << 24;
			float ewf;
			std::memcpy(&ewf, &ew, 4);
			w.n0 = { lo[0], lo[1], lo[2], ewf };
			auto pack = [](const unsigned char *b) { return (unsigned)b[0] | (unsigned)b[1] << 8 | (unsigned)b[2] << 16 | (unsigned)b[3] << 24; };
			w.n1 = { (unsigned)child_base, (unsigned)prim_base, pack(meta), pack(meta + 4) };
			w.n2 = { pack(q[0]), pack(q[0] + 4), pack(q[1]), pack(q[1] + 4) };
			w.n3 = { pack(q[2]), pack(q[2] + 4), pack(q[3]), pack(q[3] + 4) };
			w.n4 = { pack(q[4]), pack(q[4] + 4), pack(q[5]), pack(q[5] + 4) };
			out.wnodes[job.wide] = w;
		}
		out.wide_depth = max_depth;
	}
};

}  // namespace

const char *validate_edges(const double u[3], const double v[3]) {
	if (near_zero(d3(u))) return "Edge vector u cannot be zero vector";
	if (near_zero(d3(v))) return "Edge vector v cannot be zero vector";
	if (near_zero(cross(d3(u), d3(v)))) return "Edge vectors u and v cannot be collinear";
	return nullptr;
}

void make_noise_tables(uint64_t seed, double *grad768, int *perm768) {
	PhiloxKey key = philox_key(seed);
	for (int i = 0; i < 256; ++i) {
		Rnd4<double> r = rnd4<double>(key, (uint32_t)i, 0u, 0u, 0x5045524Cu);
		double z = 1.0 - 2.0 * r.x, rxy = std::sqrt(std::max(0.0, 1.0 - z * z)), phi = 2.0 * 3.14159265358979323846 * r.y;
		grad768[3 * i] = rxy * std::cos(phi);
		grad768[3 * i + 1] = rxy * std::sin(phi);
		grad768[3 * i + 2] = z;
	}
	for (int a = 0; a < 3; ++a) {
		int *p = perm768 + 256 * a;
		for (int i = 0; i < 256; ++i) p[i] = i;
		for (int i = 255; i > 0; --i) {
			U4 o = philox4x32_10(key, (uint32_t)i, (uint32_t)(a + 1), 0u, 0x5045524Cu);
			int target = (int)(o.x % (uint32_t)(i + 1));
			std::swap(p[i], p[target]);
		}
	}
}

void make_cam_basis(const double pos[3], const double target[3], const double up[3], double vfov_deg, double focus_dist,
	double defocus_angle_deg, int jitter, int W, int H, CamBasis &out) {
	D3 p = d3(pos);
	D3 f = d3(target) - p;
	f = (1.0 / len(f)) * f;
	D3 r = cross(f, d3(up));
	r = (1.0 / len(r)) * r;
	D3 u = cross(r, f);
	const double scale = std::tan(vfov_deg * 0.5 * 3.14159265358979323846 / 180.0);
	const double P[3] = { p.x, p.y, p.z }, F[3] = { f.x, f.y, f.z }, R[3] = { r.x, r.y, r.z }, U[3] = { u.x, u.y, u.z };
	std::memcpy(out.pos, P, sizeof P);
	std::memcpy(out.fwd, F, sizeof F);
	std::memcpy(out.right, R, sizeof R);
	std::memcpy(out.up, U, sizeof U);
	out.sx = ((double)W / (double)H) * scale;
	out.sy = scale;
	out.focus = focus_dist;
	out.lens_r = defocus_angle_deg > 0.0 ? focus_dist * std::tan(defocus_angle_deg * 0.5 * 3.14159265358979323846 / 180.0) : 0.0;
	out.jitter = jitter;
	out.pad_ = 0;
}

void make_rt_cam(const double pos[3], const double target[3], const double up[3], double vfov_deg, int W, int H, RtCam &out) {
	// every operation in fp32 and in rt.cpp's order: forward = (target - pos).normalized(); right = forward.cross(up).normalized();
	// up' = right.cross(forward); aspect = float(w) / h; scale = tan(fov * 0.5f * M_PI / 180.f)
	volatile float px = (float)pos[0], py = (float)pos[1], pz = (float)pos[2];
	volatile float fx = (float)target[0] - px, fy = (float)target[1] - py, fz = (float)target[2] - pz;
	volatile float il = 1.0f / std::sqrt((float)(fx * fx + fy * fy + fz * fz));
	fx = fx * il; fy = fy * il; fz = fz * il;
	const float ux = (float)up[0], uy = (float)up[1], uz = (float)up[2];
	volatile float rx = fy * uz - fz * uy, ry = fz * ux - fx * uz, rz = fx * uy - fy * ux;
	il = 1.0f / std::sqrt((float)(rx * rx + ry * ry + rz * rz));
	rx = rx * il; ry = ry * il; rz = rz * il;
	out.pos[0] = px; out.pos[1] = py; out.pos[2] = pz;
	out.fwd[0] = fx; out.fwd[1] = fy; out.fwd[2] = fz;
	out.right[0] = rx; out.right[1] = ry; out.right[2] = rz;
	out.up[0] = ry * fz - rz * fy; out.up[1] = rz * fx - rx * fz; out.up[2] = rx * fy - ry * fx;
	out.aspect = (float)W / (float)H;
	out.scale = (float)std::tan((float)vfov_deg * 0.5f * 3.14159265358979323846 / 180.f);
}

bool prim_update_records(const HostPrim &p, PrimUpdate &o) {
	const float inf = std::numeric_limits<float>::infinity();
	auto down = [&](double x) { float f = (float)x; return (double)f > x ? std::nextafter(f, -inf) : f; };
	auto up = [&](double x) { float f = (float)x; return (double)f < x ? std::nextafter(f, inf) : f; };
	Box b;
	b.reset();
	const D3 Q = d3(p.Q), u = d3(p.u), v = d3(p.v);
	int kind;
	std::memset(o.rt, 0, sizeof o.rt);
	if (p.type == PT_SPHERE) {
		const double r = p.u[0];
		if (!(r > 0.0) || !std::isfinite(r)) return false;
		o.rec = sphere_form(Q, r);
		b.grow(Q - D3{ r, r, r });
		b.grow(Q + D3{ r, r, r });
		kind = HK_SPHERE;
	} else {
		if (validate_edges(p.u, p.v)) return false;
		o.rec = plane_form(Q, u, v);
		b.grow(Q); b.grow(Q + u); b.grow(Q + v);
		if (p.type == PT_QUAD) { b.grow(Q + u + v); kind = HK_QUAD; }
		else {
			kind = HK_TRI;
			const float v0[3] = { (float)p.Q[0], (float)p.Q[1], (float)p.Q[2] };
			const float e1[3] = { (float)p.u[0], (float)p.u[1], (float)p.u[2] }, e2[3] = { (float)p.v[0], (float)p.v[1], (float)p.v[2] };
			volatile float cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
			volatile float l2 = cx * cx + cy * cy + cz * cz;
			const float il = 1.0f / std::sqrt((float)l2);
			o.rt[0] = { v0[0], v0[1], v0[2], e1[0] };
			o.rt[1] = { e1[1], e1[2], e2[0], e2[1] };
			o.rt[2] = { e2[2], cx * il, cy * il, cz * il };
		}
	}
	float kindf;
	std::memcpy(&kindf, &kind, 4);
	o.lo = { down(b.lo[0]), down(b.lo[1]), down(b.lo[2]), kindf };
	o.hi = { up(b.hi[0]), up(b.hi[1]), up(b.hi[2]), 0.f };
	return true;
}

void set_build_threads(int n) { g_build_threads.store(n > 0 ? n : 0); }

bool compile_scene(const HostScene &hs, const CompileOptions &opt, CompiledScene &out, std::string &err) {
	out.reset();
	const bool verbose = getenv("ARE_CUDA_VERBOSE") != nullptr;
	auto t_phase = std::chrono::steady_clock::now();
	auto phase = [&](const char *what) {
		const auto now = std::chrono::steady_clock::now();
		if (verbose) fprintf(stderr, "[are_cuda compile] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_phase).count());
		t_phase = now;
	};
	const int nmat = (int)hs.materials.size(), ntex = (int)hs.textures.size();
	// ---- materials / textures ----
	for (const HostMaterial &m : hs.materials) {
		MaterialRec r;
		std::memset(&r, 0, sizeof r);
		r.kind = m.kind;
		for (int i = 0; i < 8; ++i) { r.p[i] = m.p[i]; r.pf[i] = (float)m.p[i]; }
		int over = -1;
		if (m.kind == MK_LAMBERTIAN || m.kind == MK_LIGHT) over = (int)m.p[0];
		else if (m.kind == MK_METAL) over = (int)m.p[1];
		if (over >= ntex) { err = "material references a texture id that does not exist"; return false; }
		out.mats.push_back(r);
	}
	for (const HostTexture &t : hs.textures) {
		TextureRec r;
		std::memset(&r, 0, sizeof r);
		r.kind = t.kind;
		r.w = t.w;
		r.h = t.h;
		for (int i = 0; i < 8; ++i) { r.p[i] = t.p[i]; r.pf[i] = (float)t.p[i]; }
		while (out.tex_data.size() % 4) out.tex_data.push_back(0.f);
		r.data_off = (long long)out.tex_data.size();
		if (t.kind == TK_IMAGE) {
			if (t.w <= 0 || t.h <= 0 || t.rgb.size() != (size_t)t.w * t.h * 3) { err = "image texture without pixel data"; return false; }
			for (double c : t.rgb) out.tex_data.push_back((float)c);
		} else if (t.kind == TK_NOISE) {
			double grad[768];
			int perm[768];
			make_noise_tables((uint64_t)t.p[1], grad, perm);
			size_t base = out.tex_data.size();
			// [256 x (gx, gy, gz, 0) fp32: one LDG.128 per lattice corner][3 x 256 int permutations][768 fp64 gradients (harness)]
			out.tex_data.resize(base + 1024 + 768 + 1536);
			for (int i = 0; i < 256; ++i) {
				for (int k = 0; k < 3; ++k) out.tex_data[base + 4 * i + k] = (float)grad[3 * i + k];
				out.tex_data[base + 4 * i + 3] = 0.f;
			}
			std::memcpy(&out.tex_data[base + 1024], perm, sizeof perm);
			std::memcpy(&out.tex_data[base + 1792], grad, sizeof grad);
		}
		out.texs.push_back(r);
	}
	// ---- device primitive order: triangles, quads, spheres ----
	// Every output position follows from the primitive's rank inside its type, so the flattening runs on all host
	// threads into pre-sized arrays (it was a serial push_back loop: 170 ms of a 1 M-primitive commit).
	std::vector<int> order(hs.prims.size());
	{
		size_t cnt[3] = { 0, 0, 0 };
		for (const HostPrim &p : hs.prims) {
			if (p.type < 0 || p.type > 2) { err = "unknown primitive type"; return false; }
			cnt[p.type]++;
		}
		size_t at[3] = { 0, cnt[0], cnt[0] + cnt[1] };
		for (size_t i = 0; i < hs.prims.size(); ++i) order[at[hs.prims[i].type]++] = (int)i;
		out.n_tri = (int)cnt[0]; out.n_quad = (int)cnt[1]; out.n_sph = (int)cnt[2];
	}
	const size_t n_tri = (size_t)out.n_tri, n_quad = (size_t)out.n_quad, n_sph = (size_t)out.n_sph;
	std::vector<Box> pbox(order.size());
	out.info.resize(order.size()); out.prim_plane.resize(order.size()); out.shade.resize(order.size());
	out.tri64.resize(9 * n_tri); out.tri_uv.resize(6 * n_tri); out.tri_uv64.resize(6 * n_tri); out.rt_tris.resize(3 * n_tri);
	out.quad64.resize(9 * n_quad); out.sph64.resize(4 * n_sph);
	std::atomic<int> bad{ 0 };
	parallel_chunks(order.size(), 8192, [&](size_t d0, size_t d1) {
	for (size_t dp = d0; dp < d1; ++dp) {
		const HostPrim &p = hs.prims[order[dp]];
		if (p.mat < 0 || p.mat >= nmat) { bad.store(1); continue; }
		if (p.tex < 0 || p.tex >= ntex) { bad.store(2); continue; }
		out.info[dp] = { order[dp], p.mat, p.tex, p.type };
		Box b;
		b.reset();
		D3 Q = d3(p.Q), u = d3(p.u), v = d3(p.v);
		if (p.type == PT_SPHERE) {
			double r = p.u[0];
			out.prim_plane[dp] = sphere_form(Q, r);
			double *dst = &out.sph64[4 * (dp - n_tri - n_quad)];
			dst[0] = p.Q[0]; dst[1] = p.Q[1]; dst[2] = p.Q[2]; dst[3] = r;
			b.grow(Q - D3{ r, r, r });
			b.grow(Q + D3{ r, r, r });
		} else {
			out.prim_plane[dp] = plane_form(Q, u, v);
			double *dst = p.type == PT_TRIANGLE ? &out.tri64[9 * dp] : &out.quad64[9 * (dp - n_tri)];
			std::memcpy(dst, p.Q, 3 * sizeof(double));
			std::memcpy(dst + 3, p.u, 3 * sizeof(double));
			std::memcpy(dst + 6, p.v, 3 * sizeof(double));
			b.grow(Q); b.grow(Q + u); b.grow(Q + v);
			if (p.type == PT_QUAD) b.grow(Q + u + v);
			else {
				for (int k = 0; k < 6; ++k) { out.tri_uv[6 * dp + k] = (float)p.uv[k]; out.tri_uv64[6 * dp + k] = p.uv[k]; }
				{  // rt.cpp:118-122 in fp32: n = (v1 - v0).cross(v2 - v0).normalized(), normalized() = v * (1.0f / length)
					const float v0[3] = { (float)p.Q[0], (float)p.Q[1], (float)p.Q[2] };
					const float e1[3] = { (float)p.u[0], (float)p.u[1], (float)p.u[2] }, e2[3] = { (float)p.v[0], (float)p.v[1], (float)p.v[2] };
					volatile float cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
					volatile float l2 = cx * cx + cy * cy + cz * cz;
					const float il = 1.0f / std::sqrt((float)l2);
					out.rt_tris[3 * dp] = { v0[0], v0[1], v0[2], e1[0] };
					out.rt_tris[3 * dp + 1] = { e1[1], e1[2], e2[0], e2[1] };
					out.rt_tris[3 * dp + 2] = { e2[2], cx * il, cy * il, cz * il };
				}
			}
		}
		pbox[dp] = b;
		{  // shading record: normal + material kind + solid colour / parameter, SHADE_FAST when that is all shading needs
			const HostMaterial &m = hs.materials[p.mat];
			int over = -1;
			if (m.kind == MK_LAMBERTIAN || m.kind == MK_LIGHT) over = (int)m.p[0];
			else if (m.kind == MK_METAL) over = (int)m.p[1];
			const int te = over >= 0 ? over : p.tex;  // the texture that colours this primitive
			const HostTexture &tx = hs.textures[te];
			const bool fast = tx.kind == TK_SOLID && m.kind != MK_REFLECTIVE;
			const double scale = m.kind == MK_LIGHT ? m.p[1] : 1.0;
			const double p0 = m.kind == MK_METAL ? m.p[0] : (m.kind == MK_DIELECTRIC ? m.p[0] : 0.0);
			const unsigned bits = (unsigned)m.kind | (fast ? (unsigned)SHADE_FAST << 8 : 0u) | ((unsigned)tx.kind & 7u) << SHADE_TEXKIND_SHIFT |
				((unsigned)std::min(te, (int)SHADE_TEXID_MASK)) << SHADE_TEXID_SHIFT;
			float bitsf;
			std::memcpy(&bitsf, &bits, 4);
			const HotPrim &pf = out.prim_plane[dp];
			ShadeRec sr;
			sr.r0 = { pf.r0.x, pf.r0.y, pf.r0.z, bitsf };
			sr.r1 = { (float)(scale * tx.p[0]), (float)(scale * tx.p[1]), (float)(scale * tx.p[2]), (float)p0 };
			if (m.kind == MK_DIELECTRIC) sr.r1 = { 1.f, 1.f, 1.f, (float)p0 };
			out.shade[dp] = sr;
		}
	}
	});
	if (bad.load() == 1) { err = "primitive references a material id that does not exist"; return false; }
	if (bad.load() == 2) { err = "primitive references a texture id that does not exist"; return false; }
	// two triangles shade identically when nothing but the hit position enters their shading
	auto same_shading = [&](int ta, int tb) {
		const PrimInfo &ia = out.info[ta], &ib = out.info[tb];
		if (ia.mat != ib.mat || ia.tex != ib.tex) return false;
		const HostMaterial &m = hs.materials[ia.mat];
		int over = -1;
		if (m.kind == MK_LAMBERTIAN || m.kind == MK_LIGHT) over = (int)m.p[0];
		else if (m.kind == MK_METAL) over = (int)m.p[1];
		const int tk = hs.textures[over >= 0 ? over : ia.tex].kind;
		if (tk != TK_SOLID && tk != TK_CHECKER_3D && tk != TK_NOISE) return false;
		const HotPrim &pa = out.prim_plane[ta], &pb = out.prim_plane[tb];
		return pa.r0.x * pb.r0.x + pa.r0.y * pb.r0.y + pa.r0.z * pb.r0.z > 0.0f;  // same orientation
	};
	phase("flatten primitives");
	// ---- hot list with parallelogram fusion ----
	HotVec hot;
	hot.reserve(order.size());
	const int nt = out.n_tri;
	std::vector<char> fused(nt, 0);
	auto push_item = [&](int kind, const HotPrim &rec, HotIds ids, const Box &b, D3 gQ, D3 gu, D3 gv) {
		HotItem it;
		it.kind = kind; it.rec = rec; it.rec2 = rec; it.ids = ids; it.box = b;
		it.gQ = gQ; it.gu = gu; it.gv = gv;
		it.dead = false;
		for (int k = 0; k < 3; ++k) it.c[k] = 0.5 * (b.lo[k] + b.hi[k]);
		hot.push_back(it);
	};
	if (opt.fuse_parallelograms && nt >= 2) {
		// Edge table as a sorted array (hash, triangle, opposite vertex) instead of a node-based hash map: 1.5 M edges of
		// a 500 k-triangle scene sort in ~0.1 s where the multimap took 1.6 s.  Equal hashes are confirmed on the exact key.
		struct EdgeUse { uint64_t h; int tri, opp; };
		auto vert = [&](int t, int k) {
			const double *q = &out.tri64[9 * (size_t)t];
			D3 Q = d3(q);
			return k == 0 ? Q : (k == 1 ? Q + d3(q + 3) : Q + d3(q + 6));
		};
		typedef std::vector<EdgeUse, DefaultInitAlloc<EdgeUse>> EdgeVec;  // filled by all threads: no serial zero-fill first
		EdgeVec edges((size_t)nt * 3);
		parallel_chunks((size_t)nt, 4096, [&](size_t t0, size_t t1) {
			for (size_t t = t0; t < t1; ++t)
				for (int k = 0; k < 3; ++k) edges[3 * t + k] = { EHash()(ekey(vert((int)t, (k + 1) % 3), vert((int)t, (k + 2) % 3))), (int)t, k };
		});
		// runs of equal hash = candidate shared edges; every (triangle, edge) remembers its run, lone edges get none
		struct Range { int first, second; };
		std::vector<Range> run_of((size_t)nt * 3, Range{ 0, 0 });
		std::atomic<int> any_run{ 0 };
		auto mark_runs = [&](size_t lo, size_t hi) {  // edges[lo, hi) sorted; runs never cross a bucket (bucket = top hash byte)
			bool found = false;
			for (size_t i = lo; i < hi;) {
				size_t j = i + 1;
				while (j < hi && edges[j].h == edges[i].h) ++j;
				if (j - i >= 2) {
					found = true;
					for (size_t m = i; m < j; ++m) run_of[3 * (size_t)edges[m].tri + edges[m].opp] = Range{ (int)i, (int)j };
				}
				i = j;
			}
			if (found) any_run.store(1);
		};
		phase("  fusion: edge hashes");
		{  // sort by (hash, triangle, edge): scatter into 256 buckets on the top hash byte, then sort the buckets in parallel
			const auto less = [](const EdgeUse &a, const EdgeUse &b) { return a.h != b.h ? a.h < b.h : (a.tri != b.tri ? a.tri < b.tri : a.opp < b.opp); };
			if (edges.size() < 65536) {
				std::sort(edges.begin(), edges.end(), less);
				mark_runs(0, edges.size());
			} else {
				// counting sort on the top hash byte with one histogram per thread: thread t's edges of bucket b go to
				// start[b] + (what threads < t hold of b), so the scatter is stable and needs no atomics
				const unsigned T = std::max(1u, std::min(host_threads(), 64u));
				std::vector<size_t> hist((size_t)T * 256, 0);
				const size_t ne = edges.size();
				{
					std::vector<std::thread> th;
					for (unsigned t = 0; t < T; ++t)
						th.emplace_back([&, t] {
							size_t *h = &hist[(size_t)t * 256];
							for (size_t i = ne * t / T; i < ne * (t + 1) / T; ++i) ++h[edges[i].h >> 56];
						});
					for (auto &x : th) x.join();
				}
				size_t start[257] = { 0 };
				for (int b = 0; b < 256; ++b) {
					size_t at = start[b];
					for (unsigned t = 0; t < T; ++t) { const size_t c = hist[(size_t)t * 256 + b]; hist[(size_t)t * 256 + b] = at; at += c; }
					start[b + 1] = at;
				}
				EdgeVec tmp(ne);
				{
					std::vector<std::thread> th;
					for (unsigned t = 0; t < T; ++t)
						th.emplace_back([&, t] {
							size_t *fill = &hist[(size_t)t * 256];
							for (size_t i = ne * t / T; i < ne * (t + 1) / T; ++i) tmp[fill[edges[i].h >> 56]++] = edges[i];
						});
					for (auto &x : th) x.join();
				}
				edges.swap(tmp);
				parallel_chunks(256, 1, [&](size_t b0, size_t b1) {
					for (size_t b = b0; b < b1; ++b) {
						std::sort(edges.begin() + start[b], edges.begin() + start[b + 1], less);
						mark_runs(start[b], start[b + 1]);
					}
				});
			}
		}
		phase("  fusion: edge sort + runs");
		if (any_run.load())  // no shared edge anywhere (a cloud of loose triangles): nothing to fuse
		for (int t = 0; t < nt; ++t) {
			if (fused[t]) continue;
			for (int k = 0; k < 3 && !fused[t]; ++k) {
				const Range range = run_of[3 * (size_t)t + k];
				if (range.second - range.first < 2) continue;  // nobody else uses this edge
				D3 d0 = vert(t, (k + 1) % 3), d1 = vert(t, (k + 2) % 3), a = vert(t, k);
				const EKey key = ekey(d0, d1);
				for (const EdgeUse *it = edges.data() + range.first; it != edges.data() + range.second; ++it) {
					int j = it->tri;
					if (j == t || fused[j]) continue;
					if (!(ekey(vert(j, (it->opp + 1) % 3), vert(j, (it->opp + 2) % 3)) == key)) continue;  // hash collision
					D3 b = vert(j, it->opp);
					D3 e = (a + b) - (d0 + d1);
					double scale = len(a - d0) + len(b - d0) + len(d1 - d0);
					if (len(e) > 1e-9 * scale) continue;
					// lower user id takes the alpha >= beta side (ties on the diagonal go to it, like a linear scan)
					int ta = t, tb = j;
					D3 va = a, vb = b;
					if (out.info[tb].user_id < out.info[ta].user_id) { std::swap(ta, tb); std::swap(va, vb); }
					HotPrim rec = plane_form(d0, va - d0, vb - d0);
					Box bx = pbox[t];
					bx.grow(pbox[j]);
					push_item(HK_QUAD, rec, { ta, same_shading(ta, tb) ? -2 - tb : tb }, bx, d0, va - d0, vb - d0);
					fused[t] = fused[j] = 1;
					out.n_fused_pairs++;
					break;
				}
			}
		}
	}
	phase("  fusion: pair search");
	{  // every primitive that was not fused becomes a hot item of its own, in device order: positions by a prefix count,
		// then all host threads fill them
		std::vector<int> pos(order.size() + 1);
		int at = (int)hot.size();
		for (size_t dp = 0; dp < order.size(); ++dp) {
			pos[dp] = at;
			at += (dp < (size_t)nt && fused[dp]) ? 0 : 1;
		}
		pos[order.size()] = at;
		hot.resize((size_t)at);
		parallel_chunks(order.size(), 8192, [&](size_t d0, size_t d1) {
			for (size_t dp = d0; dp < d1; ++dp) {
				if (pos[dp + 1] == pos[dp]) continue;  // fused triangle
				const int type = out.info[dp].type;
				const HostPrim &hp = hs.prims[order[dp]];
				HotItem &it = hot[(size_t)pos[dp]];
				it.kind = type == PT_TRIANGLE ? HK_TRI : (type == PT_QUAD ? HK_QUAD : HK_SPHERE);
				it.rec = out.prim_plane[dp]; it.rec2 = it.rec; it.ids = { (int)dp, -1 }; it.box = pbox[dp];
				it.gQ = d3(hp.Q); it.gu = d3(hp.u); it.gv = d3(hp.v);
				for (int k = 0; k < 3; ++k) it.c[k] = 0.5 * (it.box.lo[k] + it.box.hi[k]);
				it.dead = false;
			}
		});
	}
	phase("parallelogram fusion");
	// ---- box detection: parallelograms that are faces of one parallelepiped become a single slab-test primitive ----
	if (opt.fuse_boxes) {
		std::vector<int> quads;
		for (size_t i = 0; i < hot.size(); ++i)
			if (hot[i].kind == HK_QUAD) quads.push_back((int)i);
		if (quads.size() >= 4 && quads.size() <= 2048) {
			auto corners = [&](const HotItem &q, D3 c[4]) { c[0] = q.gQ; c[1] = q.gQ + q.gu; c[2] = q.gQ + q.gv; c[3] = q.gQ + q.gu + q.gv; };
			auto same_pt = [&](D3 a, D3 b, double tol) { return len(a - b) <= tol; };
			auto find_face = [&](D3 B, D3 e1, D3 e2, double tol) -> int {  // live quad whose corner set is {B, B+e1, B+e2, B+e1+e2}
				const D3 want[4] = { B, B + e1, B + e2, B + e1 + e2 };
				for (int qi : quads) {
					const HotItem &q = hot[qi];
					if (q.dead) continue;
					D3 c[4];
					corners(q, c);
					bool all = true;
					for (int a = 0; a < 4 && all; ++a) {
						bool found = false;
						for (int b = 0; b < 4; ++b) found = found || same_pt(want[a], c[b], tol);
						all = found;
					}
					if (all) return qi;
				}
				return -1;
			};
			for (int fi : quads) {
				if (hot[fi].dead) continue;
				D3 fc[4];
				corners(hot[fi], fc);
				// the four (base corner, edge, other edge) views of F
				const D3 gu = hot[fi].gu, gv = hot[fi].gv, mu = -1.0 * gu, mv = -1.0 * gv;
				const D3 bases[8] = { fc[0], fc[1], fc[2], fc[3], fc[0], fc[1], fc[2], fc[3] };
				const D3 e1s[8] = { gu, mu, gu, mu, gv, gv, mv, mv };  // the edge shared with the neighbour G
				const D3 e2s[8] = { gv, gv, mv, mv, gu, mu, gu, mu };
				bool made = false;
				for (int view = 0; view < 8 && !made; ++view) {
					const D3 B = bases[view], a = e1s[view], b = e2s[view];
					const double tol = 1e-9 * (len(a) + len(b));
					const D3 fn = cross(a, b);
					// a neighbour G sharing the edge (B, B+a) and leaving F's plane gives the third edge c
					for (int gi : quads) {
						if (gi == fi || hot[gi].dead) continue;
						D3 gc[4];
						corners(hot[gi], gc);
						int iB = -1, iA = -1;
						for (int k = 0; k < 4; ++k) {
							if (same_pt(gc[k], B, tol)) iB = k;
							if (same_pt(gc[k], B + a, tol)) iA = k;
						}
						if (iB < 0 || iA < 0) continue;
						D3 c = { 0, 0, 0 };
						bool have = false;
						for (int k = 0; k < 4 && !have; ++k) {  // the corner adjacent to B other than B+a: B+c with B+a+c also in G
							if (k == iB || k == iA) continue;
							D3 cand = gc[k] - B;
							for (int m = 0; m < 4; ++m)
								if (m != k && m != iB && m != iA && same_pt(gc[m], B + a + cand, tol)) { c = cand; have = true; }
						}
						if (!have) continue;
						if (std::fabs(dot(fn, c)) <= 1e-9 * len(fn) * len(c)) continue;  // coplanar neighbour: not a box
						const double tol3 = 1e-9 * (len(a) + len(b) + len(c));
						// faces: 0/1 = -a/+a (spanned by b,c), 2/3 = -b/+b (a,c), 4/5 = -c/+c (a,b)
						const int face[6] = { find_face(B, b, c, tol3), find_face(B + a, b, c, tol3), find_face(B, a, c, tol3),
							find_face(B + b, a, c, tol3), find_face(B, a, b, tol3), find_face(B + c, a, b, tol3) };
						int present = 0;
						bool distinct = true;
						for (int k = 0; k < 6; ++k) {
							if (face[k] >= 0) present++;
							for (int m = 0; m < k; ++m)
								if (face[k] >= 0 && face[k] == face[m]) distinct = false;
						}
						if (present < 5 || !distinct) continue;  // closed, or open on exactly one face
						const D3 edges[3] = { a, b, c };
						const D3 centre = B + 0.5 * (a + b + c);
						HotItem bx;
						bx.dead = false;
						bx.kind = HK_BOX;
						bx.ids = { -1 - out.n_boxes, -1 };
						bx.box.reset();
						// per axis: unit slab normal (pointing along +edge), centre offset, half-width, faces at -/+
						D3 nrm3[3];
						double cen[3], hw[3];
						int fminus[3], fplus[3];
						for (int i = 0; i < 3; ++i) {
							D3 n = cross(edges[(i + 1) % 3], edges[(i + 2) % 3]);
							n = (1.0 / len(n)) * n;
							if (dot(n, edges[i]) < 0) n = -1.0 * n;
							nrm3[i] = n;
							cen[i] = dot(n, centre);
							hw[i] = 0.5 * dot(n, edges[i]);
							fminus[i] = face[2 * i];
							fplus[i] = face[2 * i + 1];
						}
						// canonical form of an open box: the absent face is axis 2, "-" side (the kernel tests only that one)
						for (int i = 0; i < 3; ++i) {
							if (fminus[i] >= 0 && fplus[i] >= 0) continue;
							std::swap(nrm3[i], nrm3[2]); std::swap(cen[i], cen[2]); std::swap(hw[i], hw[2]);
							std::swap(fminus[i], fminus[2]); std::swap(fplus[i], fplus[2]);
							if (fplus[2] < 0) {  // flip the axis so that the hole is on the "-" side
								nrm3[2] = -1.0 * nrm3[2];
								cen[2] = -cen[2];
								std::swap(fminus[2], fplus[2]);
							}
							break;
						}
						f4 *slots[3] = { &bx.rec.r0, &bx.rec.r1, &bx.rec.r2 };
						for (int i = 0; i < 3; ++i) {
							*slots[i] = { (float)nrm3[i].x, (float)nrm3[i].y, (float)nrm3[i].z, (float)cen[i] };
							for (int side = 0; side < 2; ++side) {
								const int f = side ? fplus[i] : fminus[i];
								if (f >= 0) {
									out.box_faces.push_back(hot[f].ids);
									bx.box.grow(hot[f].box);
									hot[f].dead = true;
								} else out.box_faces.push_back({ -1, -1 });
							}
						}
						bx.rec2.r0 = { (float)hw[0], (float)hw[1], (float)hw[2], present == 6 ? 0.0f : 1.0f };  // .w != 0: open at face 4
						bx.rec2.r1 = { (float)(1.0 / hw[0]), (float)(1.0 / hw[1]), (float)(1.0 / hw[2]), 0 };
						bx.rec2.r2 = { 0, 0, 0, 0 };
						for (int k = 0; k < 3; ++k) bx.c[k] = 0.5 * (bx.box.lo[k] + bx.box.hi[k]);
						bx.gQ = B; bx.gu = a; bx.gv = b;
						hot.push_back(bx);
						out.n_boxes++;
						made = true;
						break;
					}
				}
			}
			hot.erase(std::remove_if(hot.begin(), hot.end(), [](const HotItem &it) { return it.dead; }), hot.end());
		}
	}
	out.n_hot = 0;
	for (const HotItem &it : hot) out.n_hot += it.kind == HK_BOX ? 2 : 1;
	phase("box detection");
	// ---- brute list (type-sorted) ----
	if (out.n_hot <= opt.brute_max) {
		int cnt[HK_KINDS] = { 0, 0, 0, 0 };
		// boxes first, open ones (a face absent: the Cornell room) before closed ones — the lean kernel compiles the
		// open-face logic out of the closed-box tests (DevScene::lean_n_open)
		out.lean_n_open = 0;
		for (int open = 1; open >= 0; --open)
			for (const HotItem &it : hot)
				if (it.kind == HK_BOX && (it.rec2.r0.w != 0.0f) == (open == 1)) { emit_item(it, out.brute, out.brute_ids); cnt[HK_BOX]++; out.lean_n_open += open; }
		for (int kind = HK_BOX + 1; kind < HK_KINDS; ++kind)
			for (const HotItem &it : hot)
				if (it.kind == kind) { emit_item(it, out.brute, out.brute_ids); cnt[kind]++; }
		out.brute_range = { 0, cnt[HK_QUAD], cnt[HK_TRI], cnt[HK_SPHERE], cnt[HK_BOX] };
	}
	// ---- lean tables (small flat-shaded scenes: hit -> shading record with no owner resolution) ----
	out.lean_ok = false;
	out.lean_shade.clear();
	out.lean_sbase.clear();
	if (!out.brute.empty() && out.n_sph == 0 && out.brute_range.ns == 0 && out.brute_range.nb <= LEAN_MAX && out.brute_range.nq <= LEAN_MAX &&
		out.brute_range.nt <= LEAN_MAX) {
		bool ok = true;
		const ShadeRec none = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } };
		auto add = [&](HotIds id) {
			if (id.a < 0) { out.lean_shade.push_back(none); return; }  // absent box face: never hit
			const ShadeRec &sr = out.shade[id.a];
			unsigned bits;
			std::memcpy(&bits, &sr.r0.w, 4);
			if (id.b >= 0 || !((bits >> 8) & 1)) ok = false;  // halves shade differently, or shading needs textures / a lobe choice
			out.lean_shade.push_back(sr);
		};
		out.lean_sbase.assign(out.brute.size(), 0);
		size_t slot = 0;
		for (int b = 0; b < out.brute_range.nb; ++b, slot += 2) {
			out.lean_sbase[slot] = out.lean_sbase[slot + 1] = (int)out.lean_shade.size();
			const int box = -1 - out.brute_ids[slot].a;
			for (int f = 0; f < 6; ++f) add(out.box_faces[6 * (size_t)box + f]);
		}
		for (; slot < out.brute.size(); ++slot) {
			out.lean_sbase[slot] = (int)out.lean_shade.size();
			add(out.brute_ids[slot]);
		}
		out.lean_ok = ok;
		if (!ok) { out.lean_shade.clear(); out.lean_sbase.clear(); }
	}
	// ---- BVH ----
	const auto bvh_t0 = std::chrono::steady_clock::now();
	if (hot.empty() || !opt.device_bvh) {  // no device-builder input this time (reset() leaves these arrays alone)
		out.lb_lo.clear(); out.lb_hi.clear(); out.lb_prims.clear(); out.lb_ids.clear(); out.lb_slot.clear();
	}
	if (!hot.empty() && opt.device_bvh) {
		// input of the device builder: conservative fp32 bounds (rounded outwards) + the records in item order
		const size_t n = hot.size();
		out.lb_lo.resize(n); out.lb_hi.resize(n); out.lb_slot.resize(n);
		const float inf = std::numeric_limits<float>::infinity();
		int slots = 0;
		for (size_t i = 0; i < n; ++i) { out.lb_slot[i] = slots; slots += hot[i].kind == HK_BOX ? 2 : 1; }
		out.lb_prims.resize(slots); out.lb_ids.resize(slots);
		const unsigned T = host_threads();
		std::vector<float> tmin(3 * (size_t)T, inf), tmax(3 * (size_t)T, -inf);
		std::atomic<unsigned> next_slot{ 0 };
		parallel_chunks(n, 8192, [&](size_t i0, size_t i1) {
			const unsigned me = next_slot.fetch_add(1);  // private min / max slot of this chunk
			auto down = [&](double x) { float f = (float)x; return (double)f > x ? std::nextafter(f, -inf) : f; };
			auto up = [&](double x) { float f = (float)x; return (double)f < x ? std::nextafter(f, inf) : f; };
			for (size_t i = i0; i < i1; ++i) {
				const HotItem &it = hot[i];
				float kindf;
				const int kind = it.kind;
				std::memcpy(&kindf, &kind, 4);
				out.lb_lo[i] = { down(it.box.lo[0]), down(it.box.lo[1]), down(it.box.lo[2]), kindf };
				out.lb_hi[i] = { up(it.box.hi[0]), up(it.box.hi[1]), up(it.box.hi[2]), 0.f };
				const float c[3] = { 0.5f * (out.lb_lo[i].x + out.lb_hi[i].x), 0.5f * (out.lb_lo[i].y + out.lb_hi[i].y), 0.5f * (out.lb_lo[i].z + out.lb_hi[i].z) };
				for (int k = 0; k < 3; ++k) { tmin[3 * me + k] = std::min(tmin[3 * me + k], c[k]); tmax[3 * me + k] = std::max(tmax[3 * me + k], c[k]); }
				const int sl = out.lb_slot[i];
				out.lb_prims[sl] = it.rec; out.lb_ids[sl] = it.ids;
				if (it.kind == HK_BOX) { out.lb_prims[sl + 1] = it.rec2; out.lb_ids[sl + 1] = it.ids; }
			}
		});
		for (int k = 0; k < 3; ++k) { out.lb_cmin[k] = inf; out.lb_cmax[k] = -inf; }
		for (unsigned t = 0; t < T; ++t)
			for (int k = 0; k < 3; ++k) { out.lb_cmin[k] = std::min(out.lb_cmin[k], tmin[3 * t + k]); out.lb_cmax[k] = std::max(out.lb_cmax[k], tmax[3 * t + k]); }
		phase("device BVH input");
	} else if (!hot.empty()) {
		Builder b(hot, out);  // one hot primitive per leaf (the traversal kernel relies on it)
		ChildRef root = b.build(0, (int)hot.size(), 0, 0, 0);
		if (root.ref < 0) out.root_leaf_meta = root.ref;  // a one-primitive scene: the root IS the leaf reference
		out.bvh_depth = root.depth;
		phase("BVH build");
		// The compressed 8-wide hierarchy is an alternative traversal structure (ARE_TRAVERSAL_WIDE), measured slower than
		// BVH2 on B200 for this kernel: built for small scenes (cheap) and on request for large ones.
		if (opt.build_wide || out.nodes.size() <= 65536) WideBuilder(out, out).build();
		if (opt.build_bvh4 && !out.nodes.empty()) {
			out.nodes4.reserve(out.nodes.size() / 2 + 8);
			Bvh4Builder b4{ out, out.nodes4 };
			b4.build(0, 0);
			out.bvh4_depth = b4.depth;
			if (3 * b4.depth + 2 > ARE_BVH4_STACK) out.nodes4.clear();  // too deep for the traversal stack: BVH2 serves
		}
	}
	out.host_bvh_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - bvh_t0).count();
	phase("wide BVH collapse");
	if (verbose) fprintf(stderr, "[are_cuda compile] %zu hot items, BVH2 %zu nodes depth %d, wide %zu nodes depth %d\n", hot.size(), out.nodes.size(), out.bvh_depth, out.wnodes.size(), out.wide_depth);
	return true;
}

}  // namespace areb
