// imgcorr_ste.cuh — per-pixel arithmetic of the single-time-effect (STE) branch of CameraCalibration.correct()
// (reference paths relative to /root/reference/imgProcessor/):
//   SingleTimeEffectDetection          features/SingleTimeEffectDetection.py:23-75
//   NoiseLevelFunction.boundedFunction camera/NoiseLevelFunction.py:94-107
//   removeSinglePixels                 filters/removeSinglePixels.py:4-33
//   MaskedMovingAverage                fancytools (absent): published incremental mean  avg += (x - avg) / n,
//                                      restated in oracle/ste.py — the one ingredient whose parity is unpinned.
// __host__ __device__, float64 throughout (the reference computes this branch in float64), so that tests/host_emul can
// run the same functions on the CPU box.
#pragma once
#include "imgcorr_core.cuh"

namespace imgcorr {

struct SteConst {
    double minY, ax, ay;   // boundedFunction(x, minY, ax, ay) = max(nan_to_num(ay * sqrt(x - ax)), minY)
    double nstd;           // SingleTimeEffectDetection(nStd=...): 4 in correct() (:401), 3 in DarkCurrentMap
};

// threshold = noise_level_function(avg) * nStd   (SingleTimeEffectDetection.py:46)
IC_HD double ste_threshold(const SteConst& c, double avg) {
    double y = dmul(c.ay, sqrt(dsub(avg, c.ax)));          // negative argument -> NaN, as numpy's x ** 0.5
    if (y != y) y = 0.0;                                   // np.nan_to_num
    else if (y > DBL_MAX) y = DBL_MAX;
    else if (y < -DBL_MAX) y = -DBL_MAX;
    y = y > c.minY ? y : c.minY;                           // np.maximum (minY is a finite calibration constant)
    return dmul(y, c.nstd);
}

// ste = (image - avg) > threshold   (:60-62); NaN compares false
IC_HD bool ste_flag(double img, double avg, double thr) { return dsub(img, avg) > thr; }

// MaskedMovingAverage.update(image, clean) for one clean pixel: n += 1; avg += (image - avg) / n
IC_HD double ste_update(double img, double avg, int n_new) { return dadd(avg, ddiv(dsub(img, avg), (double)n_new)); }

// Same value, cheaper: for uint8 / uint16 / float32 exposures the numerator is float32-derived (|img - avg| <= ~7e38,
// and a non-zero difference of such values is >= 1e-61), the divisor a small integer, so the branch-free division of
// imgcorr_core.cuh applies; non-finite numerators and float64 exposures take the IEEE division.
template <bool F32RANGE>
IC_HD double ste_update_fast(double img, double avg, int n_new) {
    const double d = dsub(img, avg);
    if (F32RANGE && fabs(d) <= 1e300) return dadd(avg, ddiv_f32range(d, (double)n_new));
    return dadd(avg, ddiv(d, (double)n_new));
}

}  // namespace imgcorr
