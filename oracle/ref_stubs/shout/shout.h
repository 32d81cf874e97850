typedef struct shout shout_t;
