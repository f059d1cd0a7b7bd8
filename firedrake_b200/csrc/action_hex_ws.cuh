// Warp-specialised variant of helmholtz_action_kernel for degree 3 (N = 4), scalar spaces, atomic
// scatter.  Included from action_hex.cu inside its anonymous namespace (shares HelmParams, Tile,
// apply_first / apply_second, fast_rcp, cp_async*).
//
// Why: in the one-role kernel every warp alternates between fp64-dense phases (contractions,
// quadrature loop) and phases with no fp64 at all (index arithmetic, cp.async staging, RED scatter,
// work-queue bookkeeping); with three 168-register warps per SM sub-partition the fp64 pipe idles
// whenever the three happen to be in the second kind of phase (profiles/r02_action_cg3_n256_summary.txt:
// fp64 pipe 66 % active).  Here one CTA of 16 warps per SM splits the roles:
//   warps 0..11  (three warpgroups, setmaxnreg.inc) : element kernel only -- wait for a staged unit
//                 of 8 cells in shared memory, compute, leave the element vectors in shared memory;
//   warps 12..15 (one warpgroup, setmaxnreg.dec)    : data movement only -- claim work, stage map
//                 rows, gather x / vertex coordinates with cp.async, RED.ADD.F64 the results.
// Mover warp 12+m serves compute warps m, m+4, m+8 (the same SM sub-partition), two stages each
// (one being computed on, one being scattered / refilled), hand-over through mbarriers:
//   full[c][s]  : 32 cp.async completions (cp.async.mbarrier.arrive.noinc) + 1 arrive by mover lane 0
//   empty[c][s] : 1 arrive by compute lane 0 once the results sit in the stage buffer
// Element arithmetic, gather order (a warp-wide gather walks one dof column over 8 consecutive layers)
// and scatter are those of helmholtz_action_kernel; results agree to rounding (atomics reorder sums).

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned long long *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = smem_u32(bar);
    unsigned ok = 0;
    long long t0 = 0;
    for (unsigned spins = 0; !ok; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && (spins & 1023u) == 1023u) {
            // a lost arrival must not hang the device: give up after ~2 s
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000ll) __trap();
        }
    }
}

template <int NCOMP_, int NMOVE_, int NSTAGE_>
struct WsLayout {
    static constexpr int N = 4, ND = 64, CW = 8;
    static constexpr int US = 68;                      // cell stride of the value buffer (doubles)
    static constexpr int CS = 28;                      // cell stride of the coordinate buffer (24 + 4 used, see GEO)
    static constexpr int NCOMP = NCOMP_, NMOVE = NMOVE_, NSTAGE = NSTAGE_;
    static constexpr int CPM = NCOMP / NMOVE;          // compute warps served by one mover
    static constexpr int THREADS = (NCOMP + NMOVE) * 32;
    // register split (setmaxnreg, per thread): NCOMP / 4 compute warpgroups + one mover warpgroup share 64 K
    // (the pool is what the CTA is LAUNCHED with: 512 x 128 or 384 x 168 registers -- launch_ws checks it)
    static constexpr int REG_LAUNCH = (65536 / THREADS) / 8 * 8;          // 512 threads: 128, 384 threads: 168
    static constexpr int REG_COMPUTE = NCOMP == 12 ? 160 : (NMOVE == 8 ? 208 : 224);
    static constexpr int REG_MOVER = NCOMP == 12 ? 32 : (NMOVE == 8 ? 48 : 56);
    // per stage (bytes)
    static constexpr int XBUF = CW * US * 8;           // 4352
    static constexpr int COORD = CW * CS * 8;          // 1664
    static constexpr int ROWSET = 2 * ND * 4 + 2 * 8 * 4;   // two columns: dof row + vertex row = 576
    static constexpr int ROWS = 2 * ROWSET;            // double buffered over fills
    static constexpr int META = 16;                    // item, nvalid
    static constexpr int LANEINFO = 32 * 4;            // per mover lane: layer | src << 30 | valid << 31
    static constexpr int STAGE = XBUF + COORD + ROWS + META + LANEINFO;      // 7312
    static constexpr int TILE = CW * 64 * 8;           // 4096
    static constexpr int WARP = TILE + NSTAGE * STAGE; // 18720
    static constexpr int OFFS = NCOMP * WARP;          // off0 (64 ints) + off1 (8 ints)
    static constexpr int BARS = OFFS + (ND + 8) * 4;   // full[12][2], empty[12][2]
    static constexpr int BYTES = BARS + NCOMP * NSTAGE * 2 * 8;
};

template <bool MASS, int NCOMP, int NMOVE, int NSTAGE, bool STASH>
__global__ void __launch_bounds__((NCOMP + NMOVE) * 32, 1)
helmholtz_action_ws_kernel(const __grid_constant__ HelmParams<4> P)
{
    constexpr int N = 4;
    using L = WsLayout<NCOMP, NMOVE, NSTAGE>;
    static_assert(NCOMP % 4 == 0 && NMOVE % 4 == 0 && NCOMP % NMOVE == 0, "whole warpgroups per role");
    static_assert(NCOMP * L::REG_COMPUTE + NMOVE * L::REG_MOVER <= (NCOMP + NMOVE) * L::REG_LAUNCH,
                  "register split exceeds the pool the CTA is launched with");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *s_off0 = reinterpret_cast<int *>(smem_raw + L::OFFS);
    int *s_off1 = s_off0 + L::ND;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + L::BARS);
    // full[c][s] = bars[(c * NSTAGE + s) * 2], empty[c][s] = bars[(c * NSTAGE + s) * 2 + 1]

    if (threadIdx.x < L::ND) s_off0[threadIdx.x] = P.off0[threadIdx.x];
    if (threadIdx.x < 8) s_off1[threadIdx.x] = P.off1[threadIdx.x];
    if (threadIdx.x < L::NCOMP * L::NSTAGE) {
        mbar_init(bars + threadIdx.x * 2, 33);
        mbar_init(bars + threadIdx.x * 2 + 1, 1);
    }
    __syncthreads();

    const int cw = lane >> 2, t = lane & 3;

    if (warp >= L::NCOMP) {
        // =========================================================== mover
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(L::REG_MOVER));
        const int m = warp - L::NCOMP;
        const bool probe_nomove = P.ws_flags & 2;      // timing probe: signal only, no gather / scatter traffic
        const int ncells = P.ncols * P.nlay_items;
        const int nitems = (ncells + L::CW - 1) / L::CW;
        int cur = 0, end = 0;
        bool drained = false;
        unsigned alive = 0;

        // slot visited at step k: compute warp m + NMOVE (k % CPM), stage (k / CPM) % NSTAGE, fill number
        // k / (CPM NSTAGE)
        constexpr int CPM = L::CPM, NSLOT = L::CPM * NSTAGE;
        auto stage_of = [&](int k) -> unsigned char * {
            return smem_raw + (m + NMOVE * (k % CPM)) * L::WARP + L::TILE + ((k / CPM) % NSTAGE) * L::STAGE;
        };
        // Claim the unit that step k will gather, decode it and start the copy of its map rows
        // (at most two distinct columns per unit) into the row set the stage is not scattering from.
        // Runs one step AHEAD (cp.async groups complete in order: the rows must be older than the
        // gather they would otherwise wait behind).
        auto claim_and_rows = [&](int k, int &item, unsigned &info) {
            item = -1;
            info = 0;
            const int f = k / NSLOT, slot = k % NSLOT;
            const bool active = f == 0 || ((alive >> slot) & 1u);
            if (active && !drained) {
                if (cur < end) {
                    item = cur++;
                } else {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(P.counter, P.chunk);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= nitems) {
                        drained = true;
                    } else {
                        item = base;
                        cur = base + 1;
                        end = min(base + P.chunk, nitems);
                    }
                }
            }
            if (item >= 0) {
                const int lin = item * L::CW + cw;
                const bool valid = lin < ncells;
                unsigned ci = __umulhi((unsigned)lin, P.nlay_rcp);
                int kk = lin - (int)ci * P.nlay_items;
                if (kk >= P.nlay_items) { kk -= P.nlay_items; ci++; }
                if (!valid) { ci = 0; kk = 0; }
                const int layer = P.lay_first + P.lay_step * kk;
                const int col = P.collist ? __ldg(P.collist + ci) : (P.col0 + (int)ci);
                const int col_first = __shfl_sync(0xffffffffu, col, 0);
                const int src = (col != col_first) ? 1 : 0;
                info = (unsigned)layer | ((unsigned)src << 30) | (valid ? 0x80000000u : 0u);
                const unsigned peers = __match_any_sync(0xffffffffu, valid ? src : -1 - cw);
                const bool lead = ((__ffs(peers) - 1) >> 2) == cw;
                if (valid && lead) {
                    int *rows_new = reinterpret_cast<int *>(stage_of(k) + L::XBUF + L::COORD + (f & 1) * L::ROWSET);
                    const int *mrow = P.map0 + (long long)col * L::ND;
                    int *sm = rows_new + src * L::ND;
#pragma unroll
                    for (int j = 0; j < 4; j++) cp_async16(sm + 4 * (t + 4 * j), mrow + 4 * (t + 4 * j));
                    int *sv = rows_new + 2 * L::ND + src * 8;
                    cp_async4(sv + t, P.map1 + (long long)col * 8 + t);
                    cp_async4(sv + t + 4, P.map1 + (long long)col * 8 + t + 4);
                }
            }
            cp_async_commit();
        };

        // info = layer | column slot << 30 | valid << 31 (per lane)
        int item;
        unsigned info;
        claim_and_rows(0, item, info);
        cp_async_commit();                 // stands for the gather of step -1 in the group count
        for (int k = 0;; k++) {
            const int cl = k % CPM, s = (k / CPM) % NSTAGE, f = k / NSLOT;
            const int slot = k % NSLOT;
            const bool active = f == 0 || ((alive >> slot) & 1u);
            const int c = m + NMOVE * cl;
            unsigned char *stage = stage_of(k);
            double *xbuf = reinterpret_cast<double *>(stage);
            double *cbuf = reinterpret_cast<double *>(stage + L::XBUF);
            const int *rows_new = reinterpret_cast<const int *>(stage + L::XBUF + L::COORD + (f & 1) * L::ROWSET);
            const int *rows_old = reinterpret_cast<const int *>(stage + L::XBUF + L::COORD + ((f & 1) ^ 1) * L::ROWSET);
            int *meta = reinterpret_cast<int *>(stage + L::XBUF + L::COORD + L::ROWS);
            unsigned *laneinfo = reinterpret_cast<unsigned *>(stage + L::XBUF + L::COORD + L::ROWS + L::META);
            unsigned long long *full = bars + (c * NSTAGE + s) * 2, *empty = full + 1;

            // rows of the NEXT step's unit go out first
            int nx_item;
            unsigned nx_info;
            claim_and_rows(k + 1, nx_item, nx_info);

            // ---- results of the unit this stage held: scatter-add
            if (active && f > 0) {
                mbar_wait(empty, (unsigned)(f - 1) & 1u);
                const unsigned oinfo = laneinfo[lane];
                if ((oinfo >> 31) && !probe_nomove) {
                    const int olayer = (int)(oinfo & 0x3fffffffu);
                    const int *sm = rows_old + ((oinfo >> 30) & 1u) * L::ND;
                    const double *xr = xbuf + cw * L::US;
#pragma unroll(L::REG_MOVER >= 48 ? 4 : 2)
                    for (int j = 0; j < 16; j++) {
                        const int loc = j * 4 + t;
                        const int g = sm[loc] + s_off0[loc] * olayer;
                        atomicAdd(P.y + g, xr[loc]);
                    }
                }
            }
            // groups in flight, oldest first: rows(k), gather(k-1), rows(k+1): rows(k) must have landed
            cp_async_wait<2>();
            __syncwarp();

            // ---- gather the new unit (or tell the compute warp that the queue is empty)
            if (active) {
                if (item >= 0) {
                    if ((info >> 31) && !probe_nomove) {
                        const int layer = (int)(info & 0x3fffffffu);
                        const int src = (int)((info >> 30) & 1u);
                        const int *sv = rows_new + 2 * L::ND + src * 8;
                        double *scd = cbuf + cw * L::CS;
#pragma unroll
                        for (int i = t; i < 24; i += 4) {
                            const int v = i / 3, a = i - v * 3;
                            const int g = sv[v] + s_off1[v] * layer;
                            cp_async8(scd + i, P.coords + (long long)g * 3 + a);
                        }
                        const int *sm = rows_new + src * L::ND;
                        double *xr = xbuf + cw * L::US;
#pragma unroll(L::REG_MOVER >= 48 ? 4 : 2)
                        for (int j = 0; j < 16; j++) {
                            const int loc = j * 4 + t;
                            const int g = sm[loc] + s_off0[loc] * layer;
                            cp_async8(xr + loc, P.x + g);
                        }
                    }
                    laneinfo[lane] = info;
                    if (lane == 0) {
                        meta[0] = item;
                        meta[1] = min(L::CW, ncells - item * L::CW);
                    }
                    alive |= 1u << slot;
                } else {
                    if (lane == 0) meta[0] = -1;
                    alive &= ~(1u << slot);
                }
                cp_async_mbar_arrive_noinc(full);
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
            }
            cp_async_commit();             // gather(k) (possibly empty)
            if (alive == 0 && k >= NSLOT - 1) break;
            item = nx_item;
            info = nx_info;
        }
        cp_async_wait<0>();
        return;
    }

    // =============================================================== compute
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(L::REG_COMPUTE));
    {
        unsigned char *wbase = smem_raw + warp * L::WARP;
        Tile<N> tile(reinterpret_cast<double *>(wbase), cw, t);
        const double eta = P.xq[t];
        const double wy_alpha = P.wq[t] * P.alpha;
        const double wy_beta = P.wq[t] * P.beta;
        int s = 0;
        unsigned par = 0;
#pragma unroll 1
        for (;; s = (s + 1 == NSTAGE) ? 0 : s + 1, par ^= (s == 0)) {
            unsigned char *stage = wbase + L::TILE + s * L::STAGE;
            double *xbuf = reinterpret_cast<double *>(stage);
            double *sc = reinterpret_cast<double *>(stage + L::XBUF) + cw * L::CS;
            const int *meta = reinterpret_cast<const int *>(stage + L::XBUF + L::COORD + L::ROWS);
            unsigned long long *full = bars + (warp * NSTAGE + s) * 2, *empty = full + 1;
            mbar_wait(full, par);
            if (meta[0] < 0) break;
            const bool valid = cw < meta[1];

            // trilinear coefficients reduced at this lane's eta.  STASH (the 160-register split): only
            // A3 / A6 (used at every point) stay in registers; c2, c4, c5, c7 (cell) and A1 (lane) are needed
            // once per zeta plane and go back into the cell's slot of the coordinate buffer:
            // [c2 c4 c5 c7 | A1 of lanes 0..3, 4 apart]
            double A1[3], A3[3], A6[3], c2[3], c4[3], c5[3], c7[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                double X000 = sc[0 * 3 + a], X001 = sc[1 * 3 + a], X010 = sc[2 * 3 + a],
                       X011 = sc[3 * 3 + a], X100 = sc[4 * 3 + a], X101 = sc[5 * 3 + a],
                       X110 = sc[6 * 3 + a], X111 = sc[7 * 3 + a];
                if (!valid) {   // keep idle lanes finite: unit cube
                    X000 = 0; X001 = (a == 2); X010 = (a == 1); X011 = (a >= 1);
                    X100 = (a == 0); X101 = (a != 1); X110 = (a != 2); X111 = 1;
                }
                const double c1 = X100 - X000;
                c2[a] = X010 - X000;
                const double c3 = X001 - X000;
                c4[a] = X110 - X100 - X010 + X000;
                c5[a] = X011 - X010 - X001 + X000;
                const double c6 = X101 - X100 - X001 + X000;
                c7[a] = X111 - X110 - X101 - X011 + X100 + X010 + X001 - X000;
                A1[a] = fma(c4[a], eta, c1);
                A3[a] = fma(c5[a], eta, c3);
                A6[a] = fma(c7[a], eta, c6);
            }
            if (STASH) {
                __syncwarp();      // every lane of the cell has read the vertex coordinates
                if (t == 0) {
                    double2 *d = reinterpret_cast<double2 *>(sc);
                    d[0] = make_double2(c2[0], c2[1]);
                    d[1] = make_double2(c2[2], c4[0]);
                    d[2] = make_double2(c4[1], c4[2]);
                    d[3] = make_double2(c5[0], c5[1]);
                    d[4] = make_double2(c5[2], c7[0]);
                    d[5] = make_double2(c7[1], c7[2]);
                }
                double2 *d = reinterpret_cast<double2 *>(sc + 12 + 4 * t);
                d[0] = make_double2(A1[0], A1[1]);
                sc[12 + 4 * t + 2] = A1[2];
                __syncwarp();
            }
            double *su = xbuf + cw * L::US;
            double u[N][N];
#pragma unroll
            for (int x = 0; x < N; x++)
#pragma unroll
                for (int yy = 0; yy < N; yy++) u[x][yy] = valid ? su[(x * N + yy) * N + t] : 0.0;

            if (!(P.ws_flags & 1)) {       // (timing probe 1: skip the element arithmetic)
            double tmp[N][N], U[N][N], Vp[N][N];
            apply_first<N, false>(P.B, u, tmp);
            apply_second<N, false>(P.B, tmp, u);
            __syncwarp();
            tile.store_Z(u);
            __syncwarp();
            tile.load_Y(tmp);
            apply_second<N, false>(P.B, tmp, U);
            __syncwarp();
            tile.store_Y(U);
            __syncwarp();
            tile.load_Z(tmp);
            apply_second<N, false>(P.Dt, tmp, u);
            __syncwarp();
            tile.store_Z(u);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = 0; j < N; j++) Vp[i][j] = 0.0;
#pragma unroll 1
            for (int qz = 0; qz < N; qz++) {
                const double zeta = P.xq[qz];
                double dz[N];
#pragma unroll
                for (int j = 0; j < N; j++) dz[j] = P.DtR[qz * N + j];
                double ca[3], pb[3], qb[3];
                if (!STASH) {
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        ca[a] = fma(A6[a], zeta, A1[a]);
                        pb[a] = fma(c5[a], zeta, c2[a]);
                        qb[a] = fma(c7[a], zeta, c4[a]);
                    }
                } else {
                    const double2 *d = reinterpret_cast<const double2 *>(sc);
                    const double2 g0 = d[0], g1 = d[1], g2 = d[2], g3 = d[3], g4 = d[4], g5 = d[5];
                    const double2 a01 = *reinterpret_cast<const double2 *>(sc + 12 + 4 * t);
                    const double a2 = sc[12 + 4 * t + 2];
                    pb[0] = fma(g3.x, zeta, g0.x);     // c5 zeta + c2
                    pb[1] = fma(g3.y, zeta, g0.y);
                    pb[2] = fma(g4.x, zeta, g1.x);
                    qb[0] = fma(g4.y, zeta, g1.y);     // c7 zeta + c4
                    qb[1] = fma(g5.x, zeta, g2.x);
                    qb[2] = fma(g5.y, zeta, g2.y);
                    ca[0] = fma(A6[0], zeta, a01.x);
                    ca[1] = fma(A6[1], zeta, a01.y);
                    ca[2] = fma(A6[2], zeta, a2);
                }
                const double wyz_a = wy_alpha * P.wq[qz];
                const double wyz_b = wy_beta * P.wq[qz];
                double *trow = tile.row_Y(qz);
#pragma unroll
                for (int qx = 0; qx < N; qx++) {
                    const double xi = P.xq[qx];
                    double cb[3], cc[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        cb[a] = fma(qb[a], xi, pb[a]);
                        cc[a] = fma(A6[a], xi, A3[a]);
                    }
                    double gx = 0.0, gz = 0.0;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        gx = fma(P.Dt[qx * N + q], U[q][0], gx);
                        gz = fma(dz[q], U[qx][q], gz);
                    }
                    const double gy = trow[qx * N * N];
                    double r0[3], r1[3], r2[3];
                    r0[0] = cb[1] * cc[2] - cb[2] * cc[1];
                    r0[1] = cb[2] * cc[0] - cb[0] * cc[2];
                    r0[2] = cb[0] * cc[1] - cb[1] * cc[0];
                    r1[0] = cc[1] * ca[2] - cc[2] * ca[1];
                    r1[1] = cc[2] * ca[0] - cc[0] * ca[2];
                    r1[2] = cc[0] * ca[1] - cc[1] * ca[0];
                    r2[0] = ca[1] * cb[2] - ca[2] * cb[1];
                    r2[1] = ca[2] * cb[0] - ca[0] * cb[2];
                    r2[2] = ca[0] * cb[1] - ca[1] * cb[0];
                    const double det = ca[0] * r0[0] + ca[1] * r0[1] + ca[2] * r0[2];
                    const double adet = fabs(det);
                    const double sc3 = wyz_a * P.wq[qx] * fast_rcp(adet);
                    double h[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) h[a] = r0[a] * gx + r1[a] * gy + r2[a] * gz;
                    const double fx = sc3 * (r0[0] * h[0] + r0[1] * h[1] + r0[2] * h[2]);
                    const double fy = sc3 * (r1[0] * h[0] + r1[1] * h[1] + r1[2] * h[2]);
                    const double fz = sc3 * (r2[0] * h[0] + r2[1] * h[1] + r2[2] * h[2]);
                    trow[qx * N * N] = fy;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        Vp[q][0] = fma(P.Dt[qx * N + q], fx, Vp[q][0]);
                        Vp[qx][q] = fma(dz[q], fz, Vp[qx][q]);
                    }
                    if (MASS) Vp[qx][0] = fma(wyz_b * P.wq[qx] * adet, U[qx][0], Vp[qx][0]);
                }
#pragma unroll
                for (int x = 0; x < N; x++) {
                    const double u0 = U[x][0], v0 = Vp[x][0];
#pragma unroll
                    for (int j = 0; j < N - 1; j++) {
                        U[x][j] = U[x][j + 1];
                        Vp[x][j] = Vp[x][j + 1];
                    }
                    U[x][N - 1] = u0;
                    Vp[x][N - 1] = v0;
                }
            }
            __syncwarp();
            tile.load_Z(tmp);
            apply_second<N, true>(P.Dt, tmp, u);
            __syncwarp();
            tile.store_Z(u);
            __syncwarp();
            tile.load_Y(tmp);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = 0; j < N; j++) Vp[i][j] += tmp[i][j];
            apply_second<N, true>(P.B, Vp, tmp);
            __syncwarp();
            tile.store_Y(tmp);
            __syncwarp();
            tile.load_Z(u);
            apply_first<N, true>(P.B, u, tmp);
            apply_second<N, true>(P.B, tmp, u);
            }
            // element vector -> stage buffer (layout Z, the positions the values were read from)
#pragma unroll
            for (int x = 0; x < N; x++)
#pragma unroll
                for (int yy = 0; yy < N; yy++) su[(x * N + yy) * N + t] = u[x][yy];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty);
        }
    }
}

template <bool MASS, int NCOMP, int NMOVE, int NSTAGE, bool STASH>
int launch_ws(cudaStream_t st, HelmParams<4> &P, int sm_count)
{
    using L = WsLayout<NCOMP, NMOVE, NSTAGE>;
    auto kern = helmholtz_action_ws_kernel<MASS, NCOMP, NMOVE, NSTAGE, STASH>;
    static const int probe = getenv("FDB_WS_PROBE") ? atoi(getenv("FDB_WS_PROBE")) : 0;
    P.ws_flags = probe;
    static bool configured = false;
    if (!configured) {
        // setmaxnreg.inc blocks until the CTA's pool has the registers: the pool is threads x the register
        // count the kernel was compiled to, so a build that came out below REG_LAUNCH would hang -- refuse it
        cudaFuncAttributes fa;
        FDB_CUDA(cudaFuncGetAttributes(&fa, kern));
        if (fa.numRegs < L::REG_LAUNCH) {
            fdb::set_error("warp-specialised action kernel compiled to %d registers, needs %d for its setmaxnreg split",
                           fa.numRegs, L::REG_LAUNCH);
            return 1;
        }
        FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
        configured = true;
    }
    const int ncells = P.ncols * P.nlay_items;
    const int nitems = (ncells + L::CW - 1) / L::CW;
    long long grid = sm_count;
    const long long need = (nitems + L::NCOMP - 1) / L::NCOMP;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    int chunk = P.nlay_items / L::CW;
    if (chunk < 8) chunk = 8;
    if (chunk > 64) chunk = 64;
    const long long per_mover = nitems / (grid * L::NMOVE) + 1;
    if (chunk > per_mover / 32 + 1) chunk = (int)(per_mover / 32 + 1);
    static const int chunk_env = getenv("FDB_CHUNK") ? atoi(getenv("FDB_CHUNK")) : 0;
    if (chunk_env > 0) chunk = chunk_env;
    P.chunk = chunk;
    P.nlay_rcp = (unsigned)(0x100000000ull / (unsigned long long)P.nlay_items);
    if (P.nlay_items == 1) P.nlay_rcp = 0xffffffffu;
    FDB_CUDA(cudaMemsetAsync(P.counter, 0, sizeof(int), st));
    kern<<<(int)grid, L::THREADS, L::BYTES, st>>>(P);
    FDB_LAUNCH_CHECK();
    return 0;
}
