/* shim.h -- pieces shared by the two reference-side bindings (run_phmm_gpu.c, controller_gpu.c). */
#ifndef TDG_SHIM_H
#define TDG_SHIM_H
#include "tagdust_b200.h"
struct parameters;
struct model_bag;
struct fasta;
/* the process-wide GPU context (created on first use; NULL + param->errmsg on failure) */
tdg_context* tdg_shim_context(struct parameters* param);
/* start creating the GPU context on a background thread (hides the CUDA start-up behind host set-up) */
void tdg_shim_warmup(void);
void tdg_shim_warmup_join(void);   /* call before exit() / at the end of the controller */
/* struct model_bag -> tdg_model through a small content-keyed cache; seg types come from param->read_structure */
tdg_model* tdg_shim_get_model(struct model_bag* mb, struct parameters* param);
/* same tables, scratch sized for reads up to max_len (the tables do not depend on the read length) */
tdg_model* tdg_shim_get_model_len(struct model_bag* mb, struct parameters* param, int max_len);
/* the -ref sequences of a struct fasta on the devices (cached; one reference file per run) */
tdg_refset* tdg_shim_get_refset(struct fasta* f, struct parameters* param);
#endif
