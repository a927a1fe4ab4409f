#pragma once
// placeholder, replaced below
