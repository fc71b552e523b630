//! `akaze` with the hot path on a B200 -- an OVERLAY for the reference crate, not a crate of its own:
//!
//!     cp rust/src/lib.rs rust/src/ffi.rs  <akaze-rust>/akaze/src/      (replaces akaze/src/lib.rs)
//!     cp rust/build.rs                    <akaze-rust>/akaze/build.rs  (+ `build = "build.rs"`, `links = "akaze_b200"` in Cargo.toml)
//!
//! `pub mod ops; pub mod types;` below are the reference's own, unmodified `akaze/src/ops/` and `akaze/src/types/`
//! directories: the public types (Config, EvolutionStep, Keypoint, Descriptor, Match, GrayFloatImage), the module
//! paths and RANSAC (`ops::estimate_fundamental_matrix`) stay exactly as they are, so every caller of the crate
//! (akaze-util's bins, akaze/tests/integration-test.rs) compiles unchanged. Only the two public functions of
//! akaze/src/lib.rs:167-275 change their bodies: everything from the gray image on goes through the C ABI of
//! include/akaze_b200.h. Uncompiled in this image (no Rust toolchain); the tested host layers are the Python
//! package next to this directory and include/akaze_b200.hpp.
#[macro_use]
extern crate log;

use std::ffi::CStr;
use std::os::raw::c_void;
use std::path::PathBuf;
use std::ptr;

mod ffi;
pub mod ops;
pub mod types;

use ops::estimate_fundamental_matrix::remove_outliers;
use types::evolution::{allocate_evolutions, Config, EvolutionStep};
use types::feature_match::Match;
use types::image::{GrayFloatImage, ImageFunctions};
use types::keypoint::{Descriptor, Keypoint};

struct Engine(*mut c_void);
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::akz_destroy(self.0) }
    }
}

thread_local! {
    // One engine per calling thread keeps the crate's "callable from any thread, no global state" contract without
    // contention (a context may also be shared: the library serialises calls per context, include/akaze_b200.h).
    // AKZ_KEEP_EVOLUTIONS because extract_features returns all ten images of every EvolutionStep (lib.rs:167-170).
    static ENGINE: Engine = unsafe {
        let mut ctx = ptr::null_mut();
        check(ffi::akz_create(0, 16384, 16384, 1, ffi::AKZ_KEEP_EVOLUTIONS, &mut ctx));
        Engine(ctx)
    };
}

fn check(rc: i32) {
    if rc != 0 {
        // the reference panics on errors (lib.rs:171 `unwrap()`); so do we
        let msg = unsafe { CStr::from_ptr(ffi::akz_last_error()) }.to_string_lossy().into_owned();
        panic!("akaze_b200 error {}: {}", rc, msg);
    }
}

fn to_ffi(o: Config) -> ffi::akz_config {
    ffi::akz_config {
        num_sublevels: o.num_sublevels,
        max_octave_evolution: o.max_octave_evolution,
        base_scale_offset: o.base_scale_offset,
        initial_contrast: o.initial_contrast,
        contrast_percentile: o.contrast_percentile,
        contrast_factor_num_bins: o.contrast_factor_num_bins as u64,
        derivative_factor: o.derivative_factor,
        detector_threshold: o.detector_threshold,
        descriptor_channels: o.descriptor_channels as u64,
        descriptor_pattern_size: o.descriptor_pattern_size as u64,
    }
}

/// One image of one EvolutionStep from the device (akz_features_evolution_download).
unsafe fn download(f: *const c_void, level: u32, kind: i32, w: usize, h: usize) -> GrayFloatImage {
    let mut img = GrayFloatImage::new(w, h);
    check(ffi::akz_features_evolution_download(f, level, kind, img.buffer.as_mut_ptr()));
    img
}

/// The returned Vec<EvolutionStep>: metadata from the reference's own allocate_evolutions (evolution.rs:135-161;
/// the library's level table is the same restatement and is checked against it in debug builds), images from the device.
unsafe fn download_all(f: *const c_void, width: u32, height: u32, options: Config) -> Vec<EvolutionStep> {
    let mut evolutions = allocate_evolutions(width, height, options);
    debug_assert_eq!(evolutions.len() as u32, ffi::akz_features_num_levels(f));
    for (l, e) in evolutions.iter_mut().enumerate() {
        let mut info = ffi::akz_level_info::default();
        check(ffi::akz_features_level_info(f, l as u32, &mut info));
        debug_assert_eq!(info.n_steps as usize, e.fed_tau_steps.len());
        let (w, h, l) = (info.width as usize, info.height as usize, l as u32);
        e.Lt = download(f, l, ffi::AKZ_LT, w, h);
        e.Lsmooth = download(f, l, ffi::AKZ_LSMOOTH, w, h);
        e.Lx = download(f, l, ffi::AKZ_LX, w, h);
        e.Ly = download(f, l, ffi::AKZ_LY, w, h);
        e.Lxx = download(f, l, ffi::AKZ_LXX, w, h);
        e.Lyy = download(f, l, ffi::AKZ_LYY, w, h);
        e.Lxy = download(f, l, ffi::AKZ_LXY, w, h);
        e.Ldet = download(f, l, ffi::AKZ_LDET, w, h);
        if l > 0 {
            // level 0 keeps its 0x0 Lflow / Lstep, as in the reference (the diffusion loop starts at level 1, lib.rs:78)
            e.Lflow = download(f, l, ffi::AKZ_LFLOW, w, h);
            e.Lstep = download(f, l, ffi::AKZ_LSTEP, w, h);
        }
    }
    evolutions
}

/// akaze::extract_features (reference lib.rs:167-194): decode + to_luma stay here, the rest is one FFI call.
pub fn extract_features(input_image_path: PathBuf, options: Config) -> (Vec<EvolutionStep>, Vec<Keypoint>, Vec<Descriptor>) {
    let input_image = image::open(input_image_path).unwrap();
    let gray = input_image.to_luma();
    let (w, h) = (gray.width(), gray.height());
    info!("Loaded a {} x {} image", w, h);
    let cfg = to_ffi(options);
    ENGINE.with(|engine| unsafe {
        let mut f = ptr::null_mut();
        check(ffi::akz_extract_u8(engine.0, gray.as_ptr(), w, h, w as usize, &cfg, &mut f));
        let n = ffi::akz_features_count(f) as usize;
        let dlen = ffi::akz_features_descriptor_len(f) as usize;
        let (keypoints, descriptors) = if n == 0 {
            (vec![], vec![])
        } else {
            let kps = std::slice::from_raw_parts(ffi::akz_features_keypoints(f), n);
            let desc = std::slice::from_raw_parts(ffi::akz_features_descriptors(f), n * ffi::AKZ_DESCRIPTOR_STRIDE);
            (
                kps.iter()
                    .map(|k| Keypoint {
                        point: (k.x, k.y),
                        response: k.response,
                        size: k.size,
                        octave: k.octave as usize,
                        class_id: k.class_id as usize,
                        angle: k.angle,
                    })
                    .collect(),
                (0..n)
                    .map(|i| Descriptor { vector: desc[i * ffi::AKZ_DESCRIPTOR_STRIDE..i * ffi::AKZ_DESCRIPTOR_STRIDE + dlen].to_vec() })
                    .collect(),
            )
        };
        let evolutions = download_all(f, w, h, options);
        ffi::akz_features_free(f);
        info!("Extracted {} features.", n);
        (evolutions, keypoints, descriptors)
    })
}

/// ops::feature_matching::descriptor_match (reference feature_matching.rs:23-94) on the GPU; any distance_threshold.
pub fn descriptor_match_b200(descriptors_0: &[Descriptor], descriptors_1: &[Descriptor], distance_threshold: usize, lowes_ratio: f64) -> Vec<Match> {
    let pack = |d: &[Descriptor]| -> Vec<u8> {
        let mut v = vec![0u8; d.len() * ffi::AKZ_DESCRIPTOR_STRIDE];
        for (i, x) in d.iter().enumerate() {
            v[i * ffi::AKZ_DESCRIPTOR_STRIDE..i * ffi::AKZ_DESCRIPTOR_STRIDE + x.vector.len()].copy_from_slice(&x.vector);
        }
        v
    };
    let (a, b) = (pack(descriptors_0), pack(descriptors_1));
    let dlen = descriptors_0.first().or_else(|| descriptors_1.first()).map(|d| d.vector.len()).unwrap_or(61) as u32;
    let mut out = vec![ffi::akz_match { index_0: 0, index_1: 0, distance: 0.0 }; descriptors_0.len().max(1)];
    let mut n_out = 0u64;
    ENGINE.with(|engine| unsafe {
        check(ffi::akz_descriptor_match(
            engine.0,
            a.as_ptr(),
            descriptors_0.len() as u64,
            b.as_ptr(),
            descriptors_1.len() as u64,
            dlen,
            ffi::AKZ_DESCRIPTOR_STRIDE,
            distance_threshold as u64,
            lowes_ratio,
            out.as_mut_ptr(),
            &mut n_out,
        ));
    });
    out[..n_out as usize]
        .iter()
        .map(|m| Match { index_0: m.index_0 as usize, index_1: m.index_1 as usize, distance: m.distance })
        .collect()
}

/// akaze::match_features (reference lib.rs:252-275): GPU descriptor_match, the reference's own host RANSAC.
pub fn match_features(
    keypoints_0: &[Keypoint],
    descriptors_0: &[Descriptor],
    keypoints_1: &[Keypoint],
    descriptors_1: &[Descriptor],
    lowes_ratio: f64,
    ransac_trials: usize,
    ransac_epsilon_inliers: f32,
) -> Vec<Match> {
    let output = descriptor_match_b200(descriptors_0, descriptors_1, 10000, lowes_ratio); // lib.rs:261-266
    remove_outliers(&keypoints_0, &keypoints_1, &output, ransac_trials, 0.05, ransac_epsilon_inliers)
}
