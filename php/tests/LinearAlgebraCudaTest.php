<?php
namespace RindowTest\Math\Matrix\LinearAlgebraCudaTest;

use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\MatlibCuda;
// the reference's test classes are not autoloaded (cf. LinearAlgebraPHPModeTest.php:12-14)
if(!class_exists('RindowTest\Math\Matrix\LinearAlgebraTest\LinearAlgebraTest')) {
    require_once __DIR__.'/LinearAlgebraTest.php';
}
use RindowTest\Math\Matrix\LinearAlgebraTest\LinearAlgebraTest as ORGTest;

/**
 * Re-runs the reference's entire LinearAlgebraTest over the B200 driver, the same way
 * LinearAlgebraPHPModeTest (tests/RindowTest/Math/Matrix/LinearAlgebraPHPModeTest.php:17-26) pins
 * the pure-PHP driver: only setUp() changes.  Drop into tests/RindowTest/Math/Matrix/.
 */
class LinearAlgebraCudaTest extends ORGTest
{
    static protected $speedtest = false;

    public function setUp() : void
    {
        $this->service = new MatlibCuda();
        if($this->service->serviceLevel()<Service::LV_ADVANCED) {
            $this->markTestSkipped('librindow_b200.so / B200 not available');
        }
    }
}
