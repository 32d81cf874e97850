typedef struct fftwf_plan_s* fftwf_plan; typedef float fftwf_complex[2];
