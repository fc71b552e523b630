//! Raw bindings to include/akaze_b200.h (one declaration per C entry point the crate uses).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct akz_config {
    pub num_sublevels: u32,
    pub max_octave_evolution: u32,
    pub base_scale_offset: f64,
    pub initial_contrast: f64,
    pub contrast_percentile: f64,
    pub contrast_factor_num_bins: u64,
    pub derivative_factor: f64,
    pub detector_threshold: f64,
    pub descriptor_channels: u64,
    pub descriptor_pattern_size: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct akz_keypoint {
    pub x: f32,
    pub y: f32,
    pub response: f32,
    pub size: f32,
    pub octave: u32,
    pub class_id: u32,
    pub angle: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct akz_match {
    pub index_0: u64,
    pub index_1: u64,
    pub distance: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct akz_level_info {
    pub octave: u32,
    pub sublevel: u32,
    pub sigma_size: u32,
    pub width: u32,
    pub height: u32,
    pub n_steps: u32,
    pub esigma: f64,
    pub etime: f64,
}

pub const AKZ_KEEP_EVOLUTIONS: u32 = 1;
pub const AKZ_DESCRIPTOR_STRIDE: usize = 64;
// enum akz_image_kind
pub const AKZ_LT: c_int = 0;
pub const AKZ_LSMOOTH: c_int = 1;
pub const AKZ_LX: c_int = 2;
pub const AKZ_LY: c_int = 3;
pub const AKZ_LXX: c_int = 4;
pub const AKZ_LYY: c_int = 5;
pub const AKZ_LXY: c_int = 6;
pub const AKZ_LFLOW: c_int = 7;
pub const AKZ_LSTEP: c_int = 8;
pub const AKZ_LDET: c_int = 9;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct akz_top2 {
    pub best_idx: u32,
    pub best: u16,
    pub second: u16,
}

extern "C" {
    pub fn akz_last_error() -> *const c_char;
    pub fn akz_create(device: c_int, max_w: u32, max_h: u32, max_batch: u32, flags: u32, out: *mut *mut c_void) -> c_int;
    pub fn akz_destroy(ctx: *mut c_void);
    pub fn akz_extract_u8(ctx: *mut c_void, gray: *const u8, w: u32, h: u32, stride: usize, cfg: *const akz_config,
                          out: *mut *mut c_void) -> c_int;
    pub fn akz_features_count(f: *const c_void) -> u64;
    pub fn akz_features_keypoints(f: *const c_void) -> *const akz_keypoint;
    pub fn akz_features_descriptors(f: *const c_void) -> *const u8;
    pub fn akz_features_descriptor_len(f: *const c_void) -> u32;
    pub fn akz_features_num_levels(f: *const c_void) -> u32;
    pub fn akz_features_level_info(f: *const c_void, level: u32, out: *mut akz_level_info) -> c_int;
    pub fn akz_features_fed_tau(f: *const c_void, level: u32, out: *mut f64, cap: u32) -> c_int;
    pub fn akz_features_evolution_download(f: *const c_void, level: u32, kind: c_int, dst: *mut f32) -> c_int;
    pub fn akz_features_free(f: *mut c_void);
    pub fn akz_descriptor_match(ctx: *mut c_void, d0: *const u8, n0: u64, d1: *const u8, n1: u64, desc_len: u32,
                                stride: usize, distance_threshold: u64, lowes_ratio: f64, out: *mut akz_match,
                                n_out: *mut u64) -> c_int;
    pub fn akz_match_top2(ctx: *mut c_void, q: *const u8, nq: u64, db: *const u8, ndb: u64, desc_len: u32, stride: usize,
                          out: *mut akz_top2) -> c_int;
    // multi-GPU matching: one context per device, NCCL all-gather + merge inside the library
    pub fn akz_context_comm_init_all(ctxs: *const *mut c_void, n: c_int) -> c_int;
    pub fn akz_match_top2_sharded(ctxs: *const *mut c_void, n_gpu: c_int, q: *const u8, nq: u64, db: *const u8, ndb: u64,
                                  desc_len: u32, stride: usize, out: *mut akz_top2) -> c_int;
}
