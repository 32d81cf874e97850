typedef struct lame_global_struct* lame_t;
