/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under dust_b200/ may include, link or call this.
 *
 * CPU restatement (plain C) of the reference's `.vox` ingest: crates/vox/src/{loader.rs,
 * collector.rs, geometry.rs} plus the part of the third-party crate `dot_vox 5.1.1` (crates.io,
 * pinned in the reference's Cargo.lock, NOT vendored under /root/reference) that the loader
 * consumes.  dot_vox's parser is restated from the published MagicaVoxel .vox format
 * (chunks SIZE/XYZI/RGBA/nTRN/nGRP/nSHP); conventions adopted from that crate's documentation:
 *   - Voxel.i = raw colour index - 1 (saturating), so it indexes `palette` directly;
 *   - `palette` = the 256 RGBA chunk entries as stored;
 *   - Rotation byte `_r`: bits 0-1 / 2-3 = column of the non-zero entry in row 0 / row 1,
 *     bits 4,5,6 = sign of rows 0,1,2.
 * Parity status: PARITY UNPINNED for this file — the reference holds no test or fixture for
 * its loader; anchored on the call sites loader.rs:96-204, 238-308.
 */
#ifndef ORACLE_VOX_H
#define ORACLE_VOX_H
#include "oracle_vdb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OVoxModel {
    uint32_t size[3];        /* SIZE chunk: x, y, z in .vox axes (z up) */
    uint32_t n_voxels;
    uint8_t *xyzi;           /* n_voxels * 4 bytes: x, y, z, i (i already 0-based) */
} OVoxModel;

enum { O_NODE_TRANSFORM = 0, O_NODE_GROUP = 1, O_NODE_SHAPE = 2 };
typedef struct OVoxNode {
    int32_t kind;
    int32_t child;           /* transform */
    int32_t translation[3];  /* transform: frames[0] `_t`, default 0 */
    uint8_t rotation;        /* transform: frames[0] `_r`, default identity 0b0000100 */
    uint32_t n_children;     /* group */
    int32_t *children;
    int32_t model_id;        /* shape: models[0] */
} OVoxNode;

typedef struct OVoxScene {
    uint32_t n_models;
    OVoxModel *models;
    uint8_t palette[256][4];
    uint32_t n_nodes;
    OVoxNode *nodes;
} OVoxScene;

typedef struct OVoxInstance {
    uint32_t model_id;
    float obj2world[12];     /* row-major 3x4, as handed to the TLAS (tlas.rs:99-105) */
} OVoxInstance;

/* GPUVoxNode, crates/vox/src/geometry.rs:40-49 — 24 bytes */
typedef struct OGpuVoxNode {
    uint16_t x, y, z, w;
    uint64_t mask;
    uint32_t material_ptr;
    uint32_t avg_albedo;
} OGpuVoxNode;

typedef struct OVoxModelData {
    OTree *tree;             /* hierarchy!(4,2,2) */
    uint8_t size[3];         /* [size.x, size.z, size.y]  (loader.rs:292) */
    uint64_t n_materials;
    uint8_t *materials;      /* compact palette indices, 1 byte per voxel */
    uint32_t n_leaves;
    OGpuVoxNode *nodes;      /* iter_leaf order */
    float *aabbs;            /* 6 floats per leaf: min xyz, max xyz */
} OVoxModelData;

OVoxScene *o_vox_parse(const uint8_t *bytes, uint64_t len);
void       o_vox_scene_free(OVoxScene *s);
/* scene-graph walk, loader.rs:59-176; returns the number of instances (may exceed cap) */
uint32_t   o_vox_instances(const OVoxScene *s, OVoxInstance *out, uint32_t cap);
/* loader.rs:238-308 + geometry.rs:55-128 */
OVoxModelData *o_vox_load_model(const OVoxScene *s, uint32_t model_id);
void       o_vox_model_free(OVoxModelData *m);
/* loader.rs:178-204: (translation, rotation byte, size) -> row-major 3x4 local transform */
void       o_vox_to_transform(const int32_t translation[3], uint8_t rotation, const uint32_t size[3], float out[12]);

#ifdef __cplusplus
}
#endif
#endif
