namespace libconfig { class Setting; class Config; }
