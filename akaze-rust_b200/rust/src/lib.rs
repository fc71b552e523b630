//! `akaze` with the hot path on a B200: same public signatures as akaze/src/lib.rs:167-275 of the reference.
//! Uncompiled in this image (no Rust toolchain); the Python package next to it is the tested host layer.
#[macro_use]
extern crate log;

use std::ffi::CStr;
use std::os::raw::c_void;
use std::path::PathBuf;
use std::ptr;

mod ffi;
pub mod ops;   // estimate_fundamental_matrix.rs (RANSAC) is kept verbatim from the reference
pub mod types; // Config, EvolutionStep, Keypoint, Descriptor, Match, GrayFloatImage: unchanged public types

use ops::estimate_fundamental_matrix::remove_outliers;
use types::evolution::{Config, EvolutionStep};
use types::feature_match::Match;
use types::keypoint::{Descriptor, Keypoint};

thread_local! {
    // one engine per thread: contexts are single-threaded by contract (include/akaze_b200.h)
    static ENGINE: *mut c_void = unsafe {
        let mut ctx = ptr::null_mut();
        check(ffi::akz_create(0, 8192, 8192, 1, ffi::AKZ_KEEP_EVOLUTIONS, &mut ctx));
        ctx
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

/// akaze::extract_features (reference lib.rs:167-194): decode + to_luma stay here, the rest is one FFI call.
pub fn extract_features(input_image_path: PathBuf, options: Config) -> (Vec<EvolutionStep>, Vec<Keypoint>, Vec<Descriptor>) {
    let input_image = image::open(input_image_path).unwrap();
    let gray = input_image.to_luma();
    let (w, h) = (gray.width(), gray.height());
    info!("Loaded a {} x {} image", w, h);
    let cfg = to_ffi(options);
    ENGINE.with(|&ctx| unsafe {
        let mut f = ptr::null_mut();
        check(ffi::akz_extract_u8(ctx, gray.as_ptr(), w, h, w as usize, &cfg, &mut f));
        let n = ffi::akz_features_count(f) as usize;
        let kps = std::slice::from_raw_parts(ffi::akz_features_keypoints(f), n);
        let dlen = ffi::akz_features_descriptor_len(f) as usize;
        let desc = std::slice::from_raw_parts(ffi::akz_features_descriptors(f), n * 64);
        let keypoints = kps.iter().map(|k| Keypoint {
            point: (k.x, k.y), response: k.response, size: k.size,
            octave: k.octave as usize, class_id: k.class_id as usize, angle: k.angle,
        }).collect();
        let descriptors = (0..n).map(|i| Descriptor { vector: desc[i * 64..i * 64 + dlen].to_vec() }).collect();
        let evolutions = types::evolution::download_all(f); // akz_features_level_info + _fed_tau + _evolution_download
        ffi::akz_features_free(f);
        (evolutions, keypoints, descriptors)
    })
}

/// akaze::match_features (reference lib.rs:252-275): GPU descriptor_match, host RANSAC.
pub fn match_features(keypoints_0: &[Keypoint], descriptors_0: &[Descriptor], keypoints_1: &[Keypoint],
                      descriptors_1: &[Descriptor], lowes_ratio: f64, ransac_trials: usize,
                      ransac_epsilon_inliers: f32) -> Vec<Match> {
    let pack = |d: &[Descriptor]| -> Vec<u8> {
        let mut v = vec![0u8; d.len() * 64];
        for (i, x) in d.iter().enumerate() { v[i * 64..i * 64 + x.vector.len()].copy_from_slice(&x.vector); }
        v
    };
    let (a, b) = (pack(descriptors_0), pack(descriptors_1));
    let dlen = descriptors_0.first().map(|d| d.vector.len()).unwrap_or(61) as u32;
    let mut out = vec![ffi::akz_match { index_0: 0, index_1: 0, distance: 0.0 }; descriptors_0.len()];
    let mut n_out = 0u64;
    ENGINE.with(|&ctx| unsafe {
        check(ffi::akz_descriptor_match(ctx, a.as_ptr(), descriptors_0.len() as u64, b.as_ptr(), descriptors_1.len() as u64,
                                        dlen, 64, 10000, lowes_ratio, out.as_mut_ptr(), &mut n_out));
    });
    let output: Vec<Match> = out[..n_out as usize].iter()
        .map(|m| Match { index_0: m.index_0 as usize, index_1: m.index_1 as usize, distance: m.distance }).collect();
    remove_outliers(&keypoints_0, &keypoints_1, &output, ransac_trials, 0.05, ransac_epsilon_inliers)
}
