// Host-side accuracy check of the device math (auromat_b200/csrc/amt_math.cuh compiled for the
// host; the MUFU seeds are emulated by their 20-bit truncation).  Compares the fast chain
// pixel -> lat/lon/MLat/MLT/elevation with a long-double evaluation of the reference formulas
// over random ISS-like frames and prints the maximum deviation in degrees.
//   nvcc -O2 -std=c++17 -I include -o /tmp/host_math_check scripts/host_math_check.cu && /tmp/host_math_check
#include <cstdio>
#include <random>
#include "../auromat_b200/csrc/amt_math.cuh"

using namespace amt;
typedef long double ld;

static void rotz(ld a, ld M[9]) { ld c = cosl(a), s = sinl(a); ld R[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; memcpy(M, R, sizeof R); }
static void rotx(ld a, ld M[9]) { ld c = cosl(a), s = sinl(a); ld R[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(M, R, sizeof R); }
static void mm(const ld A[9], const ld B[9], ld C[9]) {
    ld T[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T[3*i+j] = 0; for (int k = 0; k < 3; ++k) T[3*i+j] += A[3*i+k] * B[3*k+j]; }
    memcpy(C, T, sizeof T);
}

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const ld PI = 3.14159265358979323846264338327950288L, K = 180.0L / PI;
    double worst[5] = {0, 0, 0, 0, 0};
    long hits = 0, total = 0;
    for (int frame = 0; frame < 400; ++frame) {
        FrameC f;
        memset(&f, 0, sizeof f);
        const int W = 4256, H = 2832;
        f.W = W; f.H = H;
        f.crpix0 = W / 2 + 1; f.crpix1 = H / 2 + 1;
        const double scale = 0.00946, th = 2 * 3.14159265 * U(rng);
        f.cd[0] = -scale * cos(th); f.cd[1] = -scale * sin(th); f.cd[2] = scale * sin(th); f.cd[3] = -scale * cos(th);
        // camera on a 400 km orbit, random position; boresight towards the limb-ish
        ld A[9], B[9], C[9], R[9];
        rotz(2 * PI * U(rng), A); rotx(PI * U(rng), B); rotz(2 * PI * U(rng), C);
        mm(A, B, R); mm(R, C, R);
        for (int k = 0; k < 9; ++k) f.rot[k] = (double)R[k];
        // camera: 6780 km along a direction ~ (60..75 deg) away from the boresight's opposite
        ld bore[3] = {R[2], R[5], R[8]};                 // rot * (0,0,1)
        ld ax[3] = {R[0], R[3], R[6]};
        const ld off = (18.0L + 8.0L * U(rng)) * PI / 180;   // nadir angle of the boresight ~ 62..72 deg
        const ld nad = (90.0L - 20.0L) * PI / 180 - off + 0.45L;
        ld down[3];
        for (int k = 0; k < 3; ++k) down[k] = cosl(nad) * bore[k] + sinl(nad) * ax[k];
        for (int k = 0; k < 3; ++k) f.cam[k] = (double)(-down[k] * 6780.0L);
        const double a = 6378.137 + 110, b = 6378.137 * (1 - 1 / 298.257223563) + 110;
        f.rad[0] = f.rad[1] = 1 / a; f.rad[2] = 1 / b;
        for (int k = 0; k < 3; ++k) f.otr[k] = -f.cam[k] * f.rad[k];
        f.oDO = f.otr[0] * f.otr[0] + f.otr[1] * f.otr[1] + f.otr[2] * f.otr[2];
        f.origin_inside = f.oDO < 1.0;
        rotz(2 * PI * U(rng), A);
        for (int k = 0; k < 9; ++k) f.m_geo[k] = (double)A[k];
        rotz(2 * PI * U(rng), A); rotx(0.3 * U(rng), B); mm(A, B, C);
        for (int k = 0; k < 9; ++k) f.m_sm[k] = (double)C[k];
        const double wa = 6378.137, wb = 6378.137 * (1 - 1 / 298.257223563);
        f.a = wa; f.b = wb; f.b_over_a = 2.0 * (wb / wa);
        f.e2a = (wa * wa - wb * wb) / (wa * wa) * wa; f.d = (wa * wa - wb * wb) / wb;
        fill_affine(f);
        for (int s = 0; s < 4000; ++s) {
            const int ix = (int)(U(rng) * W), iy = (int)(U(rng) * H);
            double dk[3], dc[3], P[3];
            dirs_kc<false>(f, nullptr, nullptr, ix, iy, dk, dc);
            bool gz;
            ++total;
            if (!intersect(f, dc, P, gz) || gz) continue;
            ++hits;
            double lat, lon, r2, mlat, mlt;
            point_to_geo(f, P, lat, lon, r2);
            point_to_mag(f, P, mlat, mlt);
            const double el = elevation_deg<false>(dc, P, r2);
            // long-double reference from the reference formulas (pixel -> everything)
            ld u = ((ld)ix - f.crpix0) + 1, v = ((ld)iy - f.crpix1) + 1;
            ld x = f.cd[0] * u + f.cd[1] * v, y = f.cd[2] * u + f.cd[3] * v;
            ld rr = sqrtl(x * x + y * y), phi = atan2l(x, -y), theta = atanl(K / rr);
            ld l = cosl(theta) * cosl(phi), m = cosl(theta) * sinl(phi), n = sinl(theta);
            ld d[3];
            for (int k = 0; k < 3; ++k) d[k] = (ld)f.rot[3*k] * l + (ld)f.rot[3*k+1] * m + (ld)f.rot[3*k+2] * n;
            ld D[3], O[3];
            for (int k = 0; k < 3; ++k) { D[k] = d[k] * f.rad[k]; O[k] = -(ld)f.cam[k] * f.rad[k]; }
            ld dDO = D[0]*O[0]+D[1]*O[1]+D[2]*O[2], dDD = D[0]*D[0]+D[1]*D[1]+D[2]*D[2], oDO = O[0]*O[0]+O[1]*O[1]+O[2]*O[2];
            ld rt = dDO * dDO - oDO * dDD + dDD;
            if (rt < 0) continue;
            ld t = (f.origin_inside ? dDO + sqrtl(rt) : dDO - sqrtl(rt)) / dDD;
            ld Q[3];
            for (int k = 0; k < 3; ++k) Q[k] = d[k] * t + f.cam[k];
            ld G[3], S[3];
            for (int k = 0; k < 3; ++k) { G[k] = 0; S[k] = 0; for (int j = 0; j < 3; ++j) { G[k] += (ld)f.m_geo[3*k+j] * Q[j]; S[k] += (ld)f.m_sm[3*k+j] * Q[j]; } }
            ld e2 = ((ld)wa * wa - (ld)wb * wb) / ((ld)wa * wa), dd = ((ld)wa * wa - (ld)wb * wb) / wb;
            ld p = sqrtl(G[0]*G[0]+G[1]*G[1]), r = sqrtl(p*p+G[2]*G[2]);
            ld tu = wb * G[2] * (1 + dd / r) / (wa * p), cu = 1 / sqrtl(1 + tu * tu), cu3 = cu*cu*cu;
            ld rlat = atanl((G[2] + dd * cu3 * tu * tu * tu) / (p - e2 * wa * cu3)) * K, rlon = atan2l(G[1], G[0]) * K;
            ld rmlat = atan2l(S[2], sqrtl(S[0]*S[0]+S[1]*S[1])) * K, rmlt = atan2l(S[1], S[0]) * K / 15 + 12;
            ld dn = sqrtl(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]), qn = sqrtl(Q[0]*Q[0]+Q[1]*Q[1]+Q[2]*Q[2]);
            ld rel = 90 - acosl(-(d[0]*Q[0]+d[1]*Q[1]+d[2]*Q[2]) / (dn * qn)) * K;
            const double e[5] = {fabs((double)(lat - rlat)), fabs((double)(lon - rlon)), fabs((double)(mlat - rmlat)),
                                 fabs((double)(mlt - rmlt)) * 15, fabs((double)(el - rel))};
            for (int k = 0; k < 5; ++k) if (e[k] > worst[k] && e[k] < 1.0) worst[k] = e[k];
            for (int k = 0; k < 5; ++k) if (e[k] >= 1.0 && e[k] < 359.0) printf("BAD k=%d e=%g lat=%g lon=%g rlon=%Lg\n", k, e[k], lat, lon, rlon);
        }
    }
    printf("rays %ld, hits %ld\nmax |delta| [deg]: lat %.3g lon %.3g mlat %.3g mlt*15 %.3g elev %.3g\n", total, hits,
           worst[0], worst[1], worst[2], worst[3], worst[4]);
    return 0;
}
