// Host-side design code (double precision, no GPU).  See host_design.h.
//
// Every routine cites the reference function whose arithmetic it reproduces;
// the operation order is kept where the reference documents it as part of its
// parity contract (window angles, mel breakpoints, bin frequencies, Kaiser
// design, planner cost search), so plans and coefficients come out identical
// to the reference's Config.create on every machine.
#include "host_design.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <limits>

namespace smb {

std::string format(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return std::string(buf);
}

static const double kPi = 3.14159265358979323846;

// ============================ windows ========================================

// I0 by its power series (window.ml:104-111): terms t_k = t_{k-1} q / k^2.
static double i0_series_window(double x) {
  const double q = 0.25 * x * x;
  if (q == 0.0) return 1.0;
  double term = 1.0, sum = 1.0;
  for (int64_t k = 1;; ++k) {
    term = term * q / double(k * k);
    sum += term;
    if (term <= 1e-17 * sum) return sum;
  }
}

void window_validate(const char* op, int kind, double param) {
  if (kind < 0 || kind >= kWindowKinds)
    throw invalid_argument(format("%s: unknown window kind %d", op, kind));
  // window.ml:77-98
  if (kind == kKaiser && !(std::isfinite(param) && param >= 0.0))
    throw invalid_argument(format(
        "%s: cannot use a kaiser window with beta %g (beta must be finite and "
        "non-negative)", op, param));
  if (kind == kGaussian && !(std::isfinite(param) && param > 0.0))
    throw invalid_argument(format(
        "%s: cannot use a gaussian window with standard deviation %g (standard "
        "deviation must be finite and positive)", op, param));
  if (kind == kTukey && !(param >= 0.0 && param <= 1.0))
    throw invalid_argument(format(
        "%s: cannot use a tukey window with taper %g (taper must lie in [0, 1])",
        op, param));
}

namespace {

struct Slots {
  std::vector<double>& buf;
  // A periodic fill describes m = n + 1 samples but owns n slots: the write
  // past the end is dropped (window.ml:135-136).
  void put(int64_t i, double v) {
    if (i < (int64_t)buf.size()) buf[(size_t)i] = v;
  }
};

// Generalised cosine window, window.ml:147-166: one cosine per sample, higher
// harmonics by the Chebyshev recurrence, first half mirrored.
void cosine_family(Slots s, const double* a, int order, int64_t m) {
  const double step = kPi / double(m - 1);
  for (int64_t i = 0; i <= (m - 1) / 2; ++i) {
    const double c = std::cos(double(2 * i - (m - 1)) * step);
    double acc = a[0] + a[1] * c;
    double prev = 1.0, cur = c;
    for (int k = 2; k < order; ++k) {
      const double t = 2.0 * c * cur - prev;
      acc = acc + a[k] * t;
      prev = cur;
      cur = t;
    }
    s.put(i, acc);
    s.put(m - 1 - i, acc);
  }
}

void fill_symmetric(Slots s, int kind, double param, int64_t m) {
  static const double hann[] = {0.5, 0.5};
  static const double hamming[] = {0.54, 0.46};
  static const double blackman[] = {0.42, 0.5, 0.08};
  static const double bharris[] = {0.35875, 0.48829, 0.14128, 0.01168};
  static const double nuttall[] = {0.3635819, 0.4891775, 0.1365995, 0.0106411};
  static const double flattop[] = {0.21557895, 0.41663158, 0.277263158,
                                   0.083578947, 0.006947368};
  const int64_t half = (m - 1) / 2;
  switch (kind) {
    case kRectangular:
      std::fill(s.buf.begin(), s.buf.end(), 1.0);
      break;
    case kHann: cosine_family(s, hann, 2, m); break;
    case kHamming: cosine_family(s, hamming, 2, m); break;
    case kBlackman: cosine_family(s, blackman, 3, m); break;
    case kBlackmanHarris: cosine_family(s, bharris, 4, m); break;
    case kNuttall: cosine_family(s, nuttall, 4, m); break;
    case kFlatTop: cosine_family(s, flattop, 5, m); break;
    case kBartlett: {                          // window.ml:170-176
      const double last = double(m - 1);
      for (int64_t i = 0; i <= half; ++i) {
        const double v = 2.0 * double(i) / last;
        s.put(i, v);
        s.put(m - 1 - i, v);
      }
      break;
    }
    case kGaussian: {                          // window.ml:180-188
      const double mid = double(m - 1) / 2.0;
      const double scale = -1.0 / (2.0 * param * param);
      for (int64_t i = 0; i <= half; ++i) {
        const double x = double(i) - mid;
        const double v = std::exp(x * x * scale);
        s.put(i, v);
        s.put(m - 1 - i, v);
      }
      break;
    }
    case kTukey: {                             // window.ml:195-207, 352-355
      if (param <= 0.0) {
        std::fill(s.buf.begin(), s.buf.end(), 1.0);
      } else if (param >= 1.0) {
        cosine_family(s, hann, 2, m);
      } else {
        const double last = double(m - 1);
        const int64_t width = (int64_t)std::floor(param * last / 2.0);
        const double step = 2.0 / param / last;
        for (int64_t i = 0; i <= width; ++i) {
          const double v = 0.5 * (1.0 + std::cos(kPi * (-1.0 + step * double(i))));
          s.put(i, v);
          s.put(m - 1 - i, v);
        }
        for (int64_t i = width + 1; i <= m - width - 2; ++i) s.put(i, 1.0);
      }
      break;
    }
    case kKaiser: {                            // window.ml:312-332
      // The reference's polynomial branches approximate this same ratio to a
      // few tens of ulp; the defining series is used here.
      const double alpha = double(m - 1) / 2.0;
      const double denom = i0_series_window(param);
      for (int64_t i = 0; i <= half; ++i) {
        const double r = (double(i) - alpha) / alpha;
        const double arg = std::max(0.0, 1.0 - r * r);
        const double v = i0_series_window(param * std::sqrt(arg)) / denom;
        s.put(i, v);
        s.put(m - 1 - i, v);
      }
      break;
    }
    default:
      throw invalid_argument("window: unknown kind");
  }
}

}  // namespace

std::vector<double> window_make(const char* op, int kind, double param,
                                bool periodic, int64_t n) {
  if (n < 1)
    throw invalid_argument(format(
        "%s: cannot make a %lld-point window (length must be at least 1)", op,
        (long long)n));
  window_validate(op, kind, param);
  std::vector<double> w((size_t)n, 0.0);
  if (n == 1) {                                // window.ml:362-364
    w[0] = 1.0;
    return w;
  }
  fill_symmetric(Slots{w}, kind, param, periodic ? n + 1 : n);
  return w;
}

// ============================ STFT grid ======================================

int64_t StftGeometry::left_width() const {      // stft.ml:132-139
  return alignment == kCentered ? fft / 2 : alignment == kLeft ? 0 : fft - 1;
}
int64_t StftGeometry::right_width() const {     // stft.ml:141-142
  return alignment == kCentered ? fft / 2 : 0;
}
int64_t StftGeometry::frames(int64_t n) const { // stft.ml:217-223
  if (n < 0)
    throw invalid_argument(format(
        "frames: cannot analyse a signal of length %lld (length must be "
        "non-negative)", (long long)n));
  if (n == 0) return 0;
  const int64_t padded = n + left_width() + right_width();
  if (padded < fft) return 0;
  return 1 + (padded - fft) / hop;
}

void stft_validate_geometry(const StftGeometry& g) {   // stft.ml:63-81
  if (g.fft < 1)
    throw invalid_argument(format(
        "create: cannot use an FFT of size %lld (fft_size must be at least 1)",
        (long long)g.fft));
  if (g.win_length < 1 || g.win_length > g.fft)
    throw invalid_argument(format(
        "create: cannot use a %lld-point window with an FFT of size %lld "
        "(win_length must lie in [1, fft_size])",
        (long long)g.win_length, (long long)g.fft));
  if (g.hop < 1)
    throw invalid_argument(format(
        "create: cannot advance frames by %lld samples (hop must be at least 1)",
        (long long)g.hop));
  if (g.alignment < 0 || g.alignment > kRight)
    throw invalid_argument("create: unknown alignment");
  if (g.pad < 0 || g.pad > kEdge) throw invalid_argument("create: unknown pad mode");
  if (g.scale < 0 || g.scale > kScalePsd)
    throw invalid_argument("create: unknown scale");
}

std::vector<double> stft_analysis_window(StftGeometry& g, int window_kind,
                                         double window_param) {
  // kDefault marks an omitted optional argument (stft.ml:68,75).
  if (g.win_length == kDefault) g.win_length = g.fft;
  if (g.hop == kDefault) g.hop = std::max<int64_t>(1, g.fft / 4);
  stft_validate_geometry(g);
  std::vector<double> coeff =
      window_make("create", window_kind, window_param, true, g.win_length);
  std::vector<double> w((size_t)g.fft, 0.0);
  const int64_t left = (g.fft - g.win_length) / 2;                      // stft.ml:97-102
  std::copy(coeff.begin(), coeff.end(), w.begin() + left);
  if (g.scale == kScaleMagnitude) {                                     // stft.ml:103-109
    double s = 0.0;
    for (double v : w) s += v;
    for (double& v : w) v /= s;
  } else if (g.scale == kScalePsd) {
    double s = 0.0;
    for (double v : w) s += v * v;
    s = std::sqrt(s);
    for (double& v : w) v /= s;
  }
  return w;
}

int64_t reflect_index(int64_t n, int64_t q) {   // stft.ml:300-305
  if (n == 1) return 0;
  const int64_t period = 2 * (n - 1);
  const int64_t r = ((q % period) + period) % period;
  return r < n ? r : period - r;
}

// ============================ mel ============================================

static const double kFsp = 200.0 / 3.0;         // convert.ml:74-80
static const double kMinLogHz = 1000.0;
static const double kMinLogMel = kMinLogHz / kFsp;
static double logstep() { return std::log(6.4) / 27.0; }

double hz_to_mel(double f, int scale) {         // convert.ml:82-91
  if (scale == kHtk) return std::log(f / 700.0 + 1.0) * (2595.0 / std::log(10.0));
  if (f < kMinLogHz) return f / kFsp;
  return std::log(f / kMinLogHz) / logstep() + kMinLogMel;
}

double mel_to_hz(double m, int scale) {         // convert.ml:93-102
  if (scale == kHtk) return (std::exp(m * (std::log(10.0) / 2595.0)) - 1.0) * 700.0;
  if (m < kMinLogMel) return m * kFsp;
  return std::exp((m - kMinLogMel) * logstep()) * kMinLogHz;
}

std::vector<double> mel_weights(int64_t n_mels, int64_t sample_rate,
                                int64_t fft_size, double f_min, double f_max,
                                int scale, int norm, double* f_max_out) {
  // mel.ml:119-160
  if (n_mels < 1)
    throw invalid_argument(format(
        "create: cannot build %lld mel bands (n_mels must be at least 1)",
        (long long)n_mels));
  if (sample_rate < 1)
    throw invalid_argument(format(
        "create: cannot use a sample rate of %lld Hz (sample_rate must be at "
        "least 1)", (long long)sample_rate));
  if (fft_size < 1)
    throw invalid_argument(format(
        "create: cannot use an FFT of size %lld (fft_size must be at least 1)",
        (long long)fft_size));
  if (!(std::isfinite(f_min) && f_min >= 0.0))
    throw invalid_argument(format(
        "create: cannot start the filterbank at %g Hz (f_min must be finite "
        "and non-negative)", f_min));
  const double nyquist = double(sample_rate) / 2.0;
  if (std::isnan(f_max) || f_max < 0.0) f_max = nyquist;     // option: default
  if (!(std::isfinite(f_max) && f_max > f_min))
    throw invalid_argument(format(
        "create: cannot span [%g, %g] Hz (f_max must be finite and greater "
        "than f_min)", f_min, f_max));
  if (f_max > nyquist)
    throw invalid_argument(format(
        "create: cannot extend the filterbank to %.17g Hz at a sample rate of "
        "%lld Hz (f_max must not exceed the Nyquist frequency %g)",
        f_max, (long long)sample_rate, nyquist));
  if (scale != kSlaney && scale != kHtk) throw invalid_argument("create: unknown mel scale");
  if (norm != kNormSlaney && norm != kNormNone) throw invalid_argument("create: unknown mel norm");
  if (f_max_out) *f_max_out = f_max;

  const int64_t bins = fft_size / 2 + 1;
  const int64_t count = n_mels + 2;
  // mel.ml:50-62: breakpoints equally spaced in mel, endpoint pinned.
  const double mel_min = hz_to_mel(f_min, scale), mel_max = hz_to_mel(f_max, scale);
  const double step = (mel_max - mel_min) / double(count - 1);
  std::vector<double> pts((size_t)count);
  for (int64_t i = 0; i < count; ++i) {
    const double mel = (i == count - 1) ? mel_max : double(i) * step + mel_min;
    pts[(size_t)i] = mel_to_hz(mel, scale);
  }
  std::vector<double> gap((size_t)(count - 1));
  for (int64_t i = 0; i + 1 < count; ++i) {
    gap[(size_t)i] = pts[(size_t)i + 1] - pts[(size_t)i];
    if (gap[(size_t)i] <= 0.0)
      throw invalid_argument(format(
          "create: cannot resolve %lld mel bands between %g and %g Hz "
          "(adjacent breakpoints collapse in double precision)",
          (long long)n_mels, f_min, f_max));
  }
  // mel.ml:39-43: bin frequency = k * (1 / (fft * (1 / sr))).
  const double bin_step = 1.0 / (double(fft_size) * (1.0 / double(sample_rate)));
  std::vector<double> w((size_t)(n_mels * bins), 0.0);
  for (int64_t m = 0; m < n_mels; ++m) {
    double peak = 0.0;
    for (int64_t k = 0; k < bins; ++k) {
      const double f = double(k) * bin_step;
      const double lower = -(pts[(size_t)m] - f) / gap[(size_t)m];
      const double upper = (pts[(size_t)m + 2] - f) / gap[(size_t)m + 1];
      const double v = std::max(0.0, std::min(lower, upper));
      w[(size_t)(m * bins + k)] = v;
      peak = std::max(peak, v);
    }
    if (peak <= 0.0)
      throw invalid_argument(format(
          "create: cannot support %lld mel bands with an FFT of size %lld (at "
          "least one filter spans no FFT bin; raise fft_size or lower n_mels)",
          (long long)n_mels, (long long)fft_size));
  }
  if (norm == kNormSlaney) {                    // mel.ml:110-117
    for (int64_t m = 0; m < n_mels; ++m) {
      const double g = 2.0 / (pts[(size_t)m + 2] - pts[(size_t)m]);
      for (int64_t k = 0; k < bins; ++k) w[(size_t)(m * bins + k)] *= g;
    }
  }
  return w;
}

// ============================ resampler ======================================

static int64_t gcd64(int64_t a, int64_t b) { return b == 0 ? a : gcd64(b, a % b); }
static int64_t ceil_pos(int64_t a, int64_t b) { return a <= 0 ? 0 : (a - 1) / b + 1; }

double kaiser_beta(double att) {                // resample.ml:105-109
  if (att > 50.0) return 0.1102 * (att - 8.7);
  if (att > 21.0) return 0.5842 * std::pow(att - 21.0, 0.4) + 0.07886 * (att - 21.0);
  return 0.0;
}

double kaiser_numtaps(double att, double width) {   // resample.ml:113-116
  const double n = std::ceil((att - 7.95) / 2.285 / (kPi * width) + 1.0);
  return std::fmod(n, 2.0) == 0.0 ? n + 1.0 : n;
}

double bessel_i0_series(double x) {             // resample.ml:126-137
  const double hx2 = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int64_t k = 1;; ++k) {
    // the reference multiplies by a tabulated 1/k^2 below k = 129
    term = k < 129 ? term * hx2 * (1.0 / double(k * k)) : term * hx2 / double(k * k);
    sum += term;
    if (term <= std::numeric_limits<double>::epsilon() * sum || k > 1000) return sum;
  }
}

std::vector<double> design_prototype(int64_t l, int64_t k, double fc, double beta) {
  // resample.ml:145-163: Kaiser-windowed sinc, right half evaluated and
  // mirrored, gain-normalised so the taps sum to l.
  const int64_t mid = k * l, n = 2 * mid + 1;
  const double i0b = bessel_i0_series(beta);
  std::vector<double> h((size_t)n, 0.0);
  for (int64_t i = mid; i < n; ++i) {
    const double z = double(i - mid);
    const double s = (i == mid) ? fc : std::sin(kPi * fc * z) / (kPi * z);
    const double r = z / double(mid);
    const double v = s * (bessel_i0_series(beta * std::sqrt(1.0 - r * r)) / i0b);
    h[(size_t)i] = v;
    h[(size_t)(n - 1 - i)] = v;
  }
  double sum = 0.0;
  for (double v : h) sum += v;
  const double gain = double(l) / sum;
  for (double& v : h) v *= gain;
  return h;
}

std::vector<double> bank_of_prototype(int64_t l, int64_t k, const std::vector<double>& h) {
  // resample.ml:169-179: row p, slot s = h[p + (2k - s) l], rows reversed so a
  // dot product walks the input window forward.
  const int64_t taps = 2 * k + 1, n = (int64_t)h.size();
  std::vector<double> b((size_t)(l * taps), 0.0);
  for (int64_t p = 0; p < l; ++p)
    for (int64_t s = 0; s < taps; ++s) {
      const int64_t idx = p + (taps - 1 - s) * l;
      if (idx < n) b[(size_t)(p * taps + s)] = h[(size_t)idx];
    }
  return b;
}

namespace {

// Frozen planner constants (resample.ml:230-343,461-476): never probed, so the
// plan is a function of the rates and the quality alone.
const double kBankBudgetBytes = 8.0 * 1024 * 1024;
const int64_t kL1EdgeBytes = 128 * 1024;
const int64_t kOlsCeilingMs = 130;
const double kRfftNs = 1.7, kIrfftNs = 1.9, kMulNs = 0.55, kCopyNs = 0.5;
const double kBlockFixedNs = 1000.0, kDotNsPerMac = 0.058, kDotNsPerOutput = 0.5;
const int64_t kGemmMinPhases = 64;

struct OlsGeom { int64_t n = 0, b = 0, delta = 0; bool ok = false; };

OlsGeom ols_geom(int64_t rate, int64_t l, int64_t m, int64_t k) {   // resample.ml:279-300
  OlsGeom g;
  const int64_t f_div = (l == 1) ? m : 1;
  const int64_t target = std::max<int64_t>(64, 10 * k);
  int64_t n = (f_div % 3 == 0) ? 3 : 1;
  while (n < target) n *= 2;
  if (!(n * 1000 <= kOlsCeilingMs * rate)) return g;
  const int64_t b = (n - 2 * k) / f_div * f_div;
  if (b < 1) return g;
  g.n = n;
  g.b = b;
  g.delta = (f_div - (3 * k % f_div)) % f_div;
  g.ok = true;
  return g;
}

double ols_cost(int64_t l, int64_t m, const OlsGeom& g) {           // resample.ml:348-368
  const double n = double(g.n);
  double block_ns;
  if (l > 1)
    block_ns = kRfftNs * n + (kCopyNs + kMulNs) * double(g.n * l / 2) +
               kIrfftNs * double(g.n * l) + kBlockFixedNs;
  else
    block_ns = kRfftNs * n + kMulNs * double(g.n / 2) +
               kCopyNs * double(m > 1 ? g.n + g.n / m : 0) +
               kIrfftNs * double(g.n / m) + kBlockFixedNs;
  const double outputs_per_block = double(g.b * l) / double(m);
  return block_ns / outputs_per_block / kDotNsPerMac;
}

double bank_cost(int64_t l, int64_t k) {                             // resample.ml:590-595
  const double macs = 2.0 * double(k) + 1.0;
  const double weighted = (l * (2 * k + 1) * 4 > kL1EdgeBytes) ? macs * 1.25 : macs;
  return weighted + kDotNsPerOutput / kDotNsPerMac;
}

bool bank_ok(int64_t len, int64_t k) {
  return double(len) * (2.0 * double(k) + 1.0) * 8.0 <= kBankBudgetBytes;
}

struct StageGeom {
  int64_t l = 1, m = 1, k = 0;
  double fc = 0.0;
  OlsGeom ols;
};

struct Cascade {
  bool found = false;
  double cost = 0.0;
  StageGeom s1, s2;
};

int64_t k_from_taps(double ntaps, int64_t l) {
  return (int64_t)std::ceil((ntaps - 1.0) / (2.0 * double(l)));
}

// The deterministic two-stage search of resample.ml:597-808.  Four candidate
// families are enumerated in the reference's order; a candidate replaces the
// incumbent only when strictly cheaper, so ties keep the earlier one.
Cascade plan_cascade(int64_t l, int64_t m, double attenuation, double passband,
                     int64_t sample_rate, double cost_bar) {
  const double att = attenuation + 20.0 * std::log10(2.0);
  const double sr = double(sample_rate);
  const int64_t target_i = sample_rate * l / m;
  const double target_f = double(target_i);
  const double f_n = 0.5 * std::min(sr, target_f);
  const double pass = passband * f_n;
  const bool near_unity = std::max(l, m) < 2 * std::min(l, m);
  Cascade best;
  auto take = [&](double cost, const StageGeom& a, const StageGeom& b) {
    if (best.found && best.cost <= cost) return;
    best.found = true;
    best.cost = cost;
    best.s1 = a;
    best.s2 = b;
  };

  if (m > l) {
    // wide integer decimator first, sharp rational stage at the lowest rate
    const int64_t fmax = std::min<int64_t>((m - 1) / l, 128);
    for (int64_t f = 2; f <= fmax; ++f) {
      const double f_mid = sr / double(f);
      const double stop1 = f_mid - f_n;
      const double width1 = (stop1 - pass) / (sr / 2.0);
      const double nt1 = kaiser_numtaps(att, width1);
      const int64_t k1 = std::max<int64_t>((int64_t)std::ceil((nt1 - 1.0) / 2.0), ceil_pos(m, l));
      const int64_t g2 = gcd64(l * f, m);
      const int64_t l2 = l * f / g2, m2 = m / g2;
      const double interp2 = f_mid * double(l2);
      const double width2 = (1.0 - passband) * f_n / (interp2 / 2.0);
      const double nt2 = kaiser_numtaps(att, width2);
      const int64_t k2 = std::max<int64_t>(1, k_from_taps(nt2, l2));
      if (bank_ok(1, k1) && bank_ok(l2, k2)) {
        StageGeom a, b;
        a.l = 1; a.m = f; a.k = k1; a.fc = (pass + stop1) / sr;
        b.l = l2; b.m = m2; b.k = k2; b.fc = (pass + f_n) / interp2;
        take(bank_cost(1, k1) * f_mid / target_f + bank_cost(l2, k2), a, b);
      }
    }
  } else {
    // sharp rational stage first at the low-rate end
    for (int64_t l1 = 2; l1 <= 16; ++l1)
      for (int64_t m1 = 1; m1 <= l1 - 1; ++m1) {
        if (!(gcd64(l1, m1) == 1 && l1 * m < l * m1)) continue;
        const double f_mid = sr * double(l1) / double(m1);
        const double interp1 = sr * double(l1);
        const double width1 = (1.0 - passband) * f_n / (interp1 / 2.0);
        const double nt1 = kaiser_numtaps(att, width1);
        const int64_t k1 = std::max<int64_t>(1, k_from_taps(nt1, l1));
        const int64_t g2 = gcd64(l * m1, m * l1);
        const int64_t l2 = l * m1 / g2, m2 = m * l1 / g2;
        const double interp2 = f_mid * double(l2);
        const double stop2 = f_mid - f_n;
        const double width2 = (stop2 - pass) / (interp2 / 2.0);
        const double nt2 = kaiser_numtaps(att, width2);
        const int64_t k2n = std::max<int64_t>(1, k_from_taps(nt2, l2));
        const int64_t k2 = l1 * ceil_pos(k2n, l1);
        if (bank_ok(l1, k1) && bank_ok(l2, k2)) {
          StageGeom a, b;
          a.l = l1; a.m = m1; a.k = k1; a.fc = (1.0 + passband) * f_n / interp1;
          b.l = l2; b.m = m2; b.k = k2; b.fc = (pass + stop2) / interp2;
          take(bank_cost(l1, k1) * f_mid / target_f + bank_cost(l2, k2), a, b);
        }
      }
  }

  if (m > l) {
    // wide rational stage first, sharp integer decimator by overlap-save
    for (int64_t f = 2; f <= 4; ++f) {
      if (!(f * l < m)) continue;
      const int64_t f_mid_i = f * target_i;
      const double f_mid = double(f_mid_i);
      const int64_t g1 = gcd64(f * l, m);
      const int64_t l1 = f * l / g1, m1 = m / g1;
      const double interp1 = sr * double(l1);
      const double stop1 = f_mid - f_n;
      const double width1 = (stop1 - pass) / (interp1 / 2.0);
      const double nt1 = kaiser_numtaps(att, width1);
      const int64_t k1 = std::max<int64_t>(std::max<int64_t>(1, k_from_taps(nt1, l1)), ceil_pos(m, l));
      const double width2 = (f_n - pass) / (f_mid / 2.0);
      const double nt2 = kaiser_numtaps(att, width2);
      const int64_t k2n = std::max<int64_t>(1, (int64_t)std::ceil((nt2 - 1.0) / 2.0));
      const int64_t k2 = l1 * ceil_pos(k2n, l1);
      const OlsGeom geom = ols_geom(f_mid_i, 1, f, k2);
      if (geom.ok && bank_ok(l1, k1) && bank_ok(1, k2)) {
        StageGeom a, b;
        a.l = l1; a.m = m1; a.k = k1; a.fc = (pass + stop1) / interp1;
        b.l = 1; b.m = f; b.k = k2; b.fc = (pass + f_n) / f_mid; b.ols = geom;
        take(bank_cost(l1, k1) * f_mid / target_f + ols_cost(1, f, geom), a, b);
      }
    }
  }

  const bool drift_class = double(std::max(l, m)) < 1.01 * double(std::min(l, m));
  const bool near_unity_ols = true;                                   // resample.ml:381
  if ((!near_unity || near_unity_ols) && !drift_class) {
    // sharp integer interpolator by overlap-save, then a wide rational stage
    for (int64_t f = 2; f <= 4; ++f) {
      if (f * m == l) continue;
      const int64_t f_mid_i = f * sample_rate;
      const double f_mid = double(f_mid_i);
      const double interp1 = f_mid;
      const double width1 = (f_n - pass) / (interp1 / 2.0);
      const double nt1 = kaiser_numtaps(att, width1);
      const int64_t k1 = std::max<int64_t>(std::max<int64_t>(1, k_from_taps(nt1, f)), ceil_pos(m, l));
      const int64_t g2 = gcd64(l, f * m);
      const int64_t l2 = l / g2, m2 = f * m / g2;
      const double interp2 = f_mid * double(l2);
      const double stop2 = f_mid - f_n;
      const double width2 = (stop2 - pass) / (interp2 / 2.0);
      const double nt2 = kaiser_numtaps(att, width2);
      const int64_t k2n = std::max<int64_t>(1, k_from_taps(nt2, l2));
      const int64_t k2 = f * ceil_pos(k2n, f);
      const OlsGeom geom = ols_geom(sample_rate, f, 1, k1);
      if (geom.ok && bank_ok(f, k1) && bank_ok(l2, k2)) {
        StageGeom a, b;
        a.l = f; a.m = 1; a.k = k1; a.fc = (pass + f_n) / interp1; a.ols = geom;
        b.l = l2; b.m = m2; b.k = k2; b.fc = (pass + stop2) / interp2;
        take(ols_cost(f, 1, geom) * f_mid / target_f + bank_cost(l2, k2), a, b);
      }
    }
  }

  if (!(best.found && best.cost < cost_bar)) best.found = false;
  return best;
}

ResampleStage make_stage(const StageGeom& g, double beta, int exec) {
  ResampleStage s;
  s.l = g.l; s.m = g.m; s.k = g.k; s.fc = g.fc; s.beta = beta;
  s.exec = exec;
  if (exec == kExecOls) { s.ols_n = g.ols.n; s.ols_b = g.ols.b; s.ols_delta = g.ols.delta; }
  s.proto = design_prototype(g.l, g.k, g.fc, beta);
  s.bank = bank_of_prototype(g.l, g.k, s.proto);
  return s;
}

std::string bytes_pretty(double bytes) {        // resample.ml:530-538
  auto scaled = [](double v, const char* unit) {
    return (v == std::floor(v)) ? format("%.0f %s", v, unit) : format("%.1f %s", v, unit);
  };
  if (bytes >= 1024.0 * 1024.0 * 1024.0) return scaled(bytes / (1024.0 * 1024.0 * 1024.0), "GB");
  if (bytes >= 1024.0 * 1024.0) return scaled(bytes / (1024.0 * 1024.0), "MB");
  return scaled(bytes / 1024.0, "KB");
}

}  // namespace

// ---- overlap-save plans -------------------------------------------------------

static void fft_pow2(std::vector<double>& re, std::vector<double>& im) {
  // in-place iterative radix-2, forward sign; plan creation only
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = -2.0 * kPi / double(len);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const double wr = std::cos(ang * double(k)), wi = std::sin(ang * double(k));
        const size_t a = i + k, b = i + k + len / 2;
        const double tr = re[b] * wr - im[b] * wi, ti = re[b] * wi + im[b] * wr;
        re[b] = re[a] - tr; im[b] = im[a] - ti;
        re[a] += tr; im[a] += ti;
      }
  }
}

static bool is_pow2(int64_t v) { return v >= 2 && (v & (v - 1)) == 0; }

int64_t OlsPlan::hi(int64_t block) const {         // resample.ml:1313-1315
  if (l > 1) return l * (block * b + n - 3 * k) - 1;
  const int64_t num = block * b + n - 3 * k - delta - 1;
  return num >= 0 ? num / m : -1;
}

int64_t OlsPlan::blocks_for(int64_t n_out) const {
  if (n_out <= 0) return 0;
  int64_t blk = 0;
  // hi is affine in the block index: jump close, then settle
  const int64_t per = l > 1 ? l * b : std::max<int64_t>(1, b / m);
  blk = std::max<int64_t>(0, (n_out - 1) / per - 2);
  while (hi(blk) < n_out - 1) ++blk;
  return blk + 1;
}

// xL stage as L branches: branch p filters the input with g_p[t] = proto[p + t L]
// (bank[p][s] = h[p + (2K - s) L], resample.ml:165-179, read against t = 2K - s).
static OlsPlan finish_polyphase(OlsPlan p, const std::vector<double>& proto) {
  const int64_t bins = p.n / 2 + 1;
  p.w = p.n;
  p.spectrum_re.assign((size_t)(bins * p.l), 0.0);
  p.spectrum_im.assign((size_t)(bins * p.l), 0.0);
  for (int64_t ph = 0; ph < p.l; ++ph) {
    std::vector<double> re((size_t)p.n, 0.0), im((size_t)p.n, 0.0);
    for (int64_t t = 0; t <= 2 * p.k; ++t) {
      const int64_t at = ph + t * p.l;
      if (at < (int64_t)proto.size()) re[(size_t)t] = proto[(size_t)at];
    }
    fft_pow2(re, im);
    for (int64_t i = 0; i < bins; ++i) {
      p.spectrum_re[(size_t)(ph * bins + i)] = re[(size_t)i] / double(p.n);
      p.spectrum_im[(size_t)(ph * bins + i)] = im[(size_t)i] / double(p.n);
    }
  }
  p.ok = true;
  return p;
}

static OlsPlan finish_ols(OlsPlan p, const std::vector<double>& proto) {
  const int64_t len = p.l > 1 ? p.n * p.l : p.n;             // resample.ml:858
  p.w = p.l > 1 ? p.n * p.l : (p.full_inverse ? p.n : p.n / p.m);
  const int64_t longest = std::max(p.n, p.w);
  if (!is_pow2(p.n) || !is_pow2(p.w) || longest > 16384 || (int64_t)proto.size() > len ||
      p.b < 1 || (p.m > 1 && !p.full_inverse && p.n % p.m != 0))
    return p;
  std::vector<double> re((size_t)len, 0.0), im((size_t)len, 0.0);
  std::copy(proto.begin(), proto.end(), re.begin());          // padded at the tail
  fft_pow2(re, im);
  // 1/M of the alias fold and 1/W of the inverse; a full-length inverse carries 1/N
  const double scale = p.full_inverse ? 1.0 / double(p.n)
                                      : (p.m > 1 ? 1.0 / double(p.m) : 1.0) / double(p.w);
  const int64_t bins = (p.l > 1 ? p.w : p.n) / 2 + 1;
  p.spectrum_re.resize((size_t)bins);
  p.spectrum_im.resize((size_t)bins);
  for (int64_t i = 0; i < bins; ++i) {
    p.spectrum_re[(size_t)i] = re[(size_t)i] * scale;
    p.spectrum_im[(size_t)i] = im[(size_t)i] * scale;
  }
  p.ok = true;
  return p;
}

OlsPlan ols_plan_for_stage(const ResampleStage& s) {
  OlsPlan p;
  if (s.exec != kExecOls || (s.l > 1 && s.m > 1)) return p;
  p.l = s.l; p.m = s.m; p.k = s.k;
  p.n = s.ols_n; p.b = s.ols_b; p.delta = s.ols_delta;
  if (s.l == 1 && s.m > 1 && 2 * s.k + s.m <= 1024) {
    // the warp-per-block kernel: block length 2048, hop a multiple of m, and the
    // reference's delta rule (resample.ml:279-300) so kept samples land on i m
    p.n = 2048;
    p.b = (2048 - 2 * s.k) / s.m * s.m;
    p.delta = (s.m - (3 * s.k % s.m)) % s.m;
    p.full_inverse = true;
  }
  if (s.m == 1 && s.l > 1 && s.l <= 8 && 2 * s.k + 1 <= 1024) {
    // interpolating stage on the same kernel: L input-rate branches, block 2048
    p.n = 2048;
    p.b = 2048 - 2 * s.k;
    p.delta = 0;
    p.polyphase = true;
    return finish_polyphase(p, s.proto);
  }
  return finish_ols(p, s.proto);
}

OlsPlan ols_plan_for_fir(const std::vector<double>& h) {
  OlsPlan p;
  p.k = ((int64_t)h.size() - 1) / 2;
  if (p.k <= 512) {
    // N = 2048 runs on the warp-per-block register-FFT kernel; the one-sample
    // shift delta makes every block's first output even, so stores pair up.
    p.n = 2048;
    p.b = 2048 - 2 * p.k;
    p.delta = p.k & 1;
    return finish_ols(p, h);
  }
  int64_t n = 64;
  while (n < 10 * p.k) n *= 2;                               // resample.ml:279-286
  if (n > 16384) n = 16384;
  p.n = n;
  p.b = n - 2 * p.k;
  p.delta = 0;
  if (p.b < n / 4) return p;                                 // filter too long for one block
  return finish_ols(p, h);
}

int64_t ResamplePlan::output_frames(int64_t n) const {   // resample.ml:1038-1051
  if (n < 0)
    throw invalid_argument(format(
        "output_frames: cannot resample a signal of length %lld (length must "
        "be non-negative)", (long long)n));
  if (n > 0 && n > INT64_MAX / l)
    throw invalid_argument(format(
        "output_frames: cannot resample a signal of length %lld (n * %lld "
        "overflows)", (long long)n, (long long)l));
  return ceil_pos(n * l, m);
}

std::string ResamplePlan::describe() const {    // resample.ml:1093-1137
  std::string q;
  switch (quality) {
    case 0: q = "fast"; break;
    case 1: q = "high"; break;
    case 2: q = "best"; break;
    default: q = format("custom(%g dB, %g)", attenuation, passband);
  }
  auto taps = [](const ResampleStage& s) {
    if (s.exec == kExecDirect) return format("%lld", (long long)(2 * s.k + 1));
    if (s.exec == kExecGemm) return format("%lld(gemm)", (long long)(2 * s.k + 1));
    return format("%lld(ols,N=%lld)", (long long)(2 * s.k * s.l + 1), (long long)(s.ols_n * s.l));
  };
  if (identity()) return format("resample(%lld Hz, identity)", (long long)sample_rate);
  if (stages.size() == 1)
    return format("resample(%lld -> %lld Hz, quality=%s, L/M=%lld/%lld, taps=%s, latency=%lld)",
                  (long long)sample_rate, (long long)target, q.c_str(), (long long)l,
                  (long long)m, taps(stages[0]).c_str(), (long long)latency);
  return format(
      "resample(%lld -> %lld Hz, quality=%s, L/M=%lld/%lld, stages=%lld/%lld:%s >> "
      "%lld/%lld:%s, latency=%lld)",
      (long long)sample_rate, (long long)target, q.c_str(), (long long)l, (long long)m,
      (long long)stages[0].l, (long long)stages[0].m, taps(stages[0]).c_str(),
      (long long)stages[1].l, (long long)stages[1].m, taps(stages[1]).c_str(),
      (long long)latency);
}

ResamplePlan resample_plan(int64_t sample_rate, int64_t target, int quality,
                           double attenuation, double passband) {
  // resample.ml:872-1019
  if (sample_rate < 1)
    throw invalid_argument(format(
        "create: cannot resample from %lld Hz (sample_rate must be at least 1)",
        (long long)sample_rate));
  if (target < 1)
    throw invalid_argument(format(
        "create: cannot resample to %lld Hz (target must be at least 1)",
        (long long)target));
  switch (quality) {                            // resample.ml:520-528
    case 0: attenuation = 100.0; passband = 0.913; break;
    case 1: attenuation = 126.0; passband = 0.913; break;
    case 2: attenuation = 175.0; passband = 0.913; break;
    case 3: break;
    default: throw invalid_argument("create: unknown quality");
  }
  if (!(std::isfinite(attenuation) && attenuation >= 40.0 && attenuation <= 200.0))
    throw invalid_argument(format(
        "create: cannot design a filter with %g dB of stop-band rejection "
        "(attenuation must be finite, in [40, 200])", attenuation));
  if (!(std::isfinite(passband) && passband >= 0.5 && passband <= 0.99))
    throw invalid_argument(format(
        "create: cannot preserve %g of the band (passband must be finite, in "
        "[0.5, 0.99])", passband));

  ResamplePlan plan;
  plan.sample_rate = sample_rate;
  plan.target = target;
  plan.quality = quality;
  plan.attenuation = attenuation;
  plan.passband = passband;
  const int64_t g = gcd64(sample_rate, target);
  const int64_t l = target / g, m = sample_rate / g;
  plan.l = l;
  plan.m = m;
  if (l == 1 && m == 1) {
    ResampleStage s;
    s.proto = {1.0};
    s.bank = {1.0};
    plan.stages.push_back(s);
    plan.latency = 0;
    return plan;
  }
  const int64_t big = std::max(l, m);
  const double width = (1.0 - passband) / double(big);
  const double ntaps = kaiser_numtaps(attenuation, width);
  const double k_f = std::ceil((ntaps - 1.0) / (2.0 * double(l)));
  const double bank_bytes = double(l) * (2.0 * k_f + 1.0) * 8.0;
  const bool single_fits = bank_bytes <= kBankBudgetBytes;
  const int64_t k_single = single_fits ? std::max<int64_t>(1, (int64_t)k_f) : 1;
  const double inf = std::numeric_limits<double>::infinity();
  const double cost_single = single_fits ? bank_cost(l, k_single) : inf;
  const bool gemm_single = single_fits && l >= kGemmMinPhases;
  OlsGeom single_ols;
  double single_ols_cost = 0.0;
  if (single_fits && ((m == 1 && l >= 2 && l <= 4) || (l == 1 && m >= 2 && m <= 4))) {
    const OlsGeom geom = ols_geom(sample_rate, l, m, k_single);
    if (geom.ok) {
      const double c = ols_cost(l, m, geom);
      if (c < cost_single) { single_ols = geom; single_ols_cost = c; }
    }
  }
  const double cost_bar = single_ols.ok ? single_ols_cost : cost_single;
  Cascade cas;
  if (!gemm_single) cas = plan_cascade(l, m, attenuation, passband, sample_rate, cost_bar);
  if (cas.found) {
    const double beta = kaiser_beta(attenuation + 20.0 * std::log10(2.0));
    plan.stages.push_back(make_stage(cas.s1, beta, cas.s1.ols.ok ? kExecOls : kExecDirect));
    plan.stages.push_back(make_stage(cas.s2, beta, cas.s2.ols.ok ? kExecOls : kExecDirect));
    plan.latency = cas.s1.k + cas.s2.k * cas.s1.m / cas.s1.l;
    return plan;
  }
  if (!single_fits) {
    const bool drift = double(big) < 1.01 * double(std::min(l, m));
    throw invalid_argument(format(
        "create: cannot resample %lld Hz to %lld Hz (%lld phases need a %s bank; "
        "the budget is %s, and no two-stage split brings it under)%s",
        (long long)sample_rate, (long long)target, (long long)l,
        bytes_pretty(bank_bytes).c_str(), bytes_pretty(kBankBudgetBytes).c_str(),
        drift ? " hint: near-unity conversion is clock-drift correction, which "
                "the fixed-ratio resampler does not do" : ""));
  }
  StageGeom sg;
  sg.l = l; sg.m = m; sg.k = k_single;
  sg.fc = (1.0 + passband) / (2.0 * double(big));
  int exec = kExecDirect;
  if (gemm_single) exec = kExecGemm;
  else if (single_ols.ok) { exec = kExecOls; sg.ols = single_ols; }
  plan.stages.push_back(make_stage(sg, kaiser_beta(attenuation), exec));
  plan.latency = k_single;
  return plan;
}

}  // namespace smb
