// Host-side (CPU, double precision) design code shared by the C ABI:
// window generation, STFT frame grid, mel filterbank weights and the
// resampler's filter design + cascade planner.  None of this touches the GPU;
// it is the native counterpart of the reference's Config.create functions
// (stft.ml:61-111, mel.ml:119-164, resample.ml:872-1019), which the reference
// also runs once per configuration on the host.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace smb {

// Precondition failure: maps to the reference's Invalid_argument.
struct invalid_argument : std::runtime_error {
  using std::runtime_error::runtime_error;
};

std::string format(const char* fmt, ...);

// ---- windows (window.ml) ---------------------------------------------------
enum WindowKind {
  kHann = 0, kHamming, kBlackman, kBlackmanHarris, kNuttall, kBartlett,
  kKaiser, kGaussian, kTukey, kFlatTop, kRectangular, kWindowKinds
};
void window_validate(const char* op, int kind, double param);
// n-point window in double; periodic = symmetric window of n+1 points with the
// last dropped.
std::vector<double> window_make(const char* op, int kind, double param,
                                bool periodic, int64_t n);

// ---- STFT configuration / frame grid (stft.ml) ------------------------------
enum Alignment { kCentered = 0, kLeft = 1, kRight = 2 };
enum PadKind { kReflect = 0, kConstant = 1, kEdge = 2 };
enum StftScale { kScaleNone = 0, kScaleMagnitude = 1, kScalePsd = 2 };

// Marks an omitted optional integer argument (OCaml ?hop / ?win_length).
static const int64_t kDefault = INT32_MIN;

struct StftGeometry {
  int64_t fft = 0, hop = 0, win_length = 0;
  int alignment = kCentered, pad = kReflect, scale = kScaleNone;
  double pad_value = 0.0;
  int64_t bins() const { return fft / 2 + 1; }
  int64_t left_width() const;
  int64_t right_width() const;
  int64_t frames(int64_t n) const;       // stft.ml:217-223
};
// Validates like Stft.Config.create and returns the analysis window (window
// centred in fft, normalised per scale).
std::vector<double> stft_analysis_window(StftGeometry& g, int window_kind,
                                         double window_param);
void stft_validate_geometry(const StftGeometry& g);
int64_t reflect_index(int64_t n, int64_t q);   // stft.ml:300-305

// ---- mel (mel.ml, convert.ml) ------------------------------------------------
enum MelScale { kSlaney = 0, kHtk = 1 };
enum MelNorm { kNormSlaney = 0, kNormNone = 1 };
double hz_to_mel(double f, int scale);
double mel_to_hz(double m, int scale);
// [n_mels x bins] row-major weights in double; f_max < 0 or NaN -> Nyquist.
std::vector<double> mel_weights(int64_t n_mels, int64_t sample_rate,
                                int64_t fft_size, double f_min, double f_max,
                                int scale, int norm, double* f_max_out);

// ---- resampler design + planner (resample.ml) --------------------------------
enum StageExec { kExecDirect = 0, kExecOls = 1, kExecGemm = 2 };
struct ResampleStage {
  int64_t l = 1, m = 1, k = 0;           // factors l/m, group delay k
  double fc = 0.0, beta = 0.0;
  int exec = kExecDirect;
  int64_t ols_n = 0, ols_b = 0, ols_delta = 0;
  std::vector<double> proto;              // 2*k*l + 1 taps
  std::vector<double> bank;               // [l][2k+1], rows reversed
};
struct ResamplePlan {
  int64_t sample_rate = 0, target = 0, l = 1, m = 1, latency = 0;
  double attenuation = 0.0, passband = 0.0;
  int quality = 1;                        // 0 fast, 1 high, 2 best, 3 custom
  std::vector<ResampleStage> stages;      // one or two
  bool identity() const { return l == 1 && m == 1 && latency == 0; }
  int64_t output_frames(int64_t n) const; // ceil(n*l/m)
  std::string describe() const;           // Config.pp string
};
double kaiser_beta(double att);
double kaiser_numtaps(double att, double width);
double bessel_i0_series(double x);
std::vector<double> design_prototype(int64_t l, int64_t k, double fc, double beta);
std::vector<double> bank_of_prototype(int64_t l, int64_t k, const std::vector<double>& h);
// Overlap-save geometry and plan spectrum of a stage (resample.ml:279-300, 856-867).
struct OlsPlan {
  bool ok = false;
  int64_t n = 0, b = 0, delta = 0, w = 0, l = 1, m = 1, k = 0;
  // true: the block is inverted at full length n and every m-th sample kept
  // (ols2048_kernel), so n / m need not be a transform length
  bool full_inverse = false;
  // true: an xL stage as its L polyphase branches at the input rate, each a plain
  // filter g_p[t] = h[p + t L] with its own n/2+1-bin spectrum (branch-major)
  bool polyphase = false;
  std::vector<double> spectrum_re, spectrum_im;   // W/2+1 (xL) or N/2+1 bins, 1/M and 1/W folded in
  int64_t blocks_for(int64_t n_out) const;         // blocks whose runs cover [0, n_out)
  int64_t hi(int64_t block) const;                 // last output block `block` completes
};
// GPU-executable OLS plan for a stage, or ok = false when the transform lengths
// are not powers of two / too long for one CTA (the direct kernel runs instead).
// Decimating stages whose filter fits take the GPU's own block length 2048
// (any m), whatever length the planner priced for the CPU.
OlsPlan ols_plan_for_stage(const ResampleStage& s);
// FIR as an L = M = 1 stage: N = 2048 up to 1025 taps, else the smallest power of
// two >= 10 K.
OlsPlan ols_plan_for_fir(const std::vector<double>& h);

ResamplePlan resample_plan(int64_t sample_rate, int64_t target, int quality,
                           double attenuation, double passband);

}  // namespace smb
