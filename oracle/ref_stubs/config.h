/* what cmake would generate from config.h.in; nothing the structures or the mixer depend on */
