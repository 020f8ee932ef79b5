(* soundml_b200.ml — the OCaml side of the drop-in: externals over
   libsoundml_b200.so and the three functions of the reference that change.

   UNVERIFIED (no OCaml toolchain in the build container).  The intent is a
   patch to soundml/lib that leaves every signature of stft.mli / mel.mli /
   resample.mli / soundml.mli untouched:

     Stft.power_spectrum   stft.ml:687-691   -> B200.power_spectrum
     Stft.transform        stft.ml:632-650   -> B200.transform
     Mel.apply             mel.ml:202-231    -> B200.mel_apply
     Soundml.mel_spectrogram soundml.ml:22-24 -> B200.mel_spectrogram (fused)
     Resample.apply        resample.ml:1913-1936 -> B200.resample_apply
     Stft.transform_range  stft.ml:652-666   -> B200.transform_range
     Stft.invert           stft.ml:937-939   -> B200.invert
     Convert.power_to_db / amplitude_to_db  convert.ml:20-56 -> B200.to_db
     Soundml.mfcc          soundml.ml:50-95  -> B200.mfcc

   Config.create stays in OCaml: the plan is built from the float64 window /
   weights the config already owns, so the two sides cannot disagree. *)

type stft_plan
type mel_plan
type resample_plan

type ('a, 'b) ba = ('a, 'b, Bigarray.c_layout) Bigarray.Array1.t

external stft_create :
  int -> int -> int -> int -> float -> (float, Bigarray.float64_elt) ba -> stft_plan
  = "soundml_b200_stft_create_bc" "soundml_b200_stft_create"

external power_spectrum_c :
  stft_plan -> (float, 'a) ba -> int -> int -> float -> (float, 'a) ba -> unit
  = "soundml_b200_power_spectrum_bc" "soundml_b200_power_spectrum"

external transform_c :
  stft_plan -> (float, 'a) ba -> int -> int -> (Complex.t, 'c) ba -> unit
  = "soundml_b200_transform"

external mel_create : int -> int -> (float, Bigarray.float64_elt) ba -> mel_plan
  = "soundml_b200_mel_create"

external mel_apply_c : mel_plan -> (float, 'a) ba -> int -> int -> (float, 'a) ba -> unit
  = "soundml_b200_mel_apply"

external mel_spectrogram_c :
  stft_plan -> mel_plan -> (float, 'a) ba -> int -> int -> float -> (float, 'a) ba -> unit
  = "soundml_b200_mel_spectrogram_bc" "soundml_b200_mel_spectrogram"

external resample_create : int -> int -> int -> float -> float -> resample_plan
  = "soundml_b200_resample_create"

external resample_apply_c :
  resample_plan -> (float, Bigarray.float32_elt) ba -> int -> int
  -> (float, Bigarray.float32_elt) ba -> unit
  = "soundml_b200_resample_apply"

external resample_apply_f64_c :
  resample_plan -> (float, Bigarray.float64_elt) ba -> int -> int
  -> (float, Bigarray.float64_elt) ba -> unit
  = "soundml_b200_resample_apply_f64"

external transform_range_c :
  stft_plan -> (float, 'a) ba -> int -> int -> int -> int -> (Complex.t, 'c) ba -> unit
  = "soundml_b200_transform_range_bc" "soundml_b200_transform_range"

(* length < 0 stands for None *)
external invert_c :
  stft_plan -> (Complex.t, 'c) ba -> int -> int -> int -> (float, 'a) ba -> unit
  = "soundml_b200_invert_bc" "soundml_b200_invert"

(* amplitude -> x -> reference -> amin -> top_db (nan = None) -> out *)
external to_db_c :
  bool -> (float, 'a) ba -> float -> float -> float -> (float, 'a) ba -> unit
  = "soundml_b200_to_db_bc" "soundml_b200_to_db"

(* lifter = 0. stands for None *)
external mfcc_c :
  stft_plan -> mel_plan -> (float, 'a) ba -> int -> int -> int -> float -> (float, 'a) ba
  -> unit
  = "soundml_b200_mfcc_bc" "soundml_b200_mfcc"

(* soundml-io: the layout pass after stub_readf (soundml_io_stubs.c:832-872) on the
   device; mode 1 = planar, 2 = downmix. *)
external ingest_layout_c :
  (float, 'a) ba -> int -> int -> int -> (float, 'a) ba -> int -> int -> unit
  = "soundml_b200_ingest_layout_bc" "soundml_b200_ingest_layout"

(* The flat storage of a contiguous tensor, shared (resample.ml:94). *)
let array1_of t = Nx_buffer.to_bigarray1 (Nx.to_buffer t)

let leading_shape t =
  let shape = Nx.shape t in
  Array.sub shape 0 (Array.length shape - 1)

let alignment_code = function `Centered -> 0 | `Left -> 1 | `Right -> 2

let pad_code = function `Reflect -> (0, 0.) | `Constant v -> (1, v) | `Edge -> (2, 0.)

(* One plan per configuration, created on first use (Config.t is immutable). *)
let stft_plan_of (c : Stft.Config.t) =
  let pad, pad_value = pad_code (Stft.Config.pad c) in
  stft_create (Stft.Config.fft_size c) (Stft.Config.hop c)
    (alignment_code (Stft.Config.alignment c))
    pad pad_value
    (array1_of (Stft.Config.analysis_window c))

(* Replacement body of Stft.power_spectrum: same checks, same result shape
   [...; bins; frames]; the arithmetic is the fused B200 kernel. *)
let power_spectrum ?(power = 2.) c x =
  let n = Nx.dim (Nx.ndim x - 1) x in
  let lead = leading_shape x in
  let batch = Array.fold_left ( * ) 1 lead in
  let frames = Stft.frames c ~n in
  let out = Nx.zeros (Nx.dtype x) (Array.append lead [|Stft.Config.bins c; frames|]) in
  if batch > 0 && frames > 0 then
    power_spectrum_c (stft_plan_of c) (array1_of (Nx.contiguous x)) batch n power
      (array1_of out) ;
  out

let mel_spectrogram stft_config mel_config ?(power = 2.) x =
  let n = Nx.dim (Nx.ndim x - 1) x in
  let lead = leading_shape x in
  let batch = Array.fold_left ( * ) 1 lead in
  let frames = Stft.frames stft_config ~n in
  let n_mels = Mel.Config.n_mels mel_config in
  let out = Nx.zeros (Nx.dtype x) (Array.append lead [|n_mels; frames|]) in
  ( if batch > 0 && frames > 0 then
      let mel =
        mel_create n_mels (Mel.Config.fft_size mel_config)
          (array1_of (Mel.filterbank Nx.float64 mel_config))
      in
      mel_spectrogram_c (stft_plan_of stft_config) mel
        (array1_of (Nx.contiguous x))
        batch n power (array1_of out) ) ;
  out

let resample_apply ~sample_rate ~target ~quality x =
  let q, att, pb =
    match quality with
    | `Fast -> (0, 0., 0.)
    | `High -> (1, 0., 0.)
    | `Best -> (2, 0., 0.)
    | `Custom (att, pb) -> (3, att, pb)
  in
  let plan = resample_create sample_rate target q att pb in
  let n = Nx.dim (Nx.ndim x - 1) x in
  let lead = leading_shape x in
  let batch = Array.fold_left ( * ) 1 lead in
  let g = let rec gcd a b = if b = 0 then a else gcd b (a mod b) in gcd sample_rate target in
  let l = target / g and m = sample_rate / g in
  let total = if n <= 0 then 0 else (((n * l) - 1) / m) + 1 in
  let out = Nx.zeros Nx.float32 (Array.append lead [|total|]) in
  if batch > 0 && n > 0 then
    resample_apply_c plan (array1_of (Nx.contiguous x)) batch n (array1_of out) ;
  out

(* Replacement body of Stft.invert: the preconditions (check_synthesis,
   stft.ml:787) stay in OCaml; the library re-validates and raises the same
   messages. *)
let invert dtype c ?length z =
  let rank = Nx.ndim z in
  let frames = Nx.dim (rank - 1) z in
  let lead = Array.sub (Nx.shape z) 0 (rank - 2) in
  let batch = Array.fold_left ( * ) 1 lead in
  let out_len =
    match length with Some n -> n | None -> Stft.output_length c ~frames
  in
  let out = Nx.zeros dtype (Array.append lead [|out_len|]) in
  if batch > 0 && out_len > 0 && frames > 0 then
    invert_c (stft_plan_of c) (array1_of (Nx.contiguous z)) batch frames
      (match length with Some n -> n | None -> -1)
      (array1_of out) ;
  out

(* Replacement body of Convert.power_to_db / amplitude_to_db: elementwise over
   the flat storage, the top_db clamp is one whole-tensor maximum on the GPU. *)
let to_db ~amplitude ?(reference = 1.) ~amin ?top_db s =
  let out = Nx.zeros (Nx.dtype s) (Nx.shape s) in
  if Nx.numel s > 0 then
    to_db_c amplitude (array1_of (Nx.contiguous s)) reference amin
      (match top_db with Some t -> t | None -> Float.nan)
      (array1_of out) ;
  out

let mfcc stft_config mel_config ?(n_mfcc = 20) ?lifter x =
  let n = Nx.dim (Nx.ndim x - 1) x in
  let lead = leading_shape x in
  let batch = Array.fold_left ( * ) 1 lead in
  let frames = Stft.frames stft_config ~n in
  let out = Nx.zeros (Nx.dtype x) (Array.append lead [|n_mfcc; frames|]) in
  ( if batch > 0 && frames > 0 then
      let mel =
        mel_create (Mel.Config.n_mels mel_config) (Mel.Config.fft_size mel_config)
          (array1_of (Mel.filterbank Nx.float64 mel_config))
      in
      mfcc_c (stft_plan_of stft_config) mel (array1_of (Nx.contiguous x)) batch n n_mfcc
        (match lifter with Some l -> l | None -> 0.)
        (array1_of out) ) ;
  out
